// Row-side building blocks shared by the tensor-core render kernels (fine_tc.cu, coarse_tc.cu): fast sin/cos for
// the bf16 PE operand, the cooperative VM gather task, and the UMMA weight-slice packer.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace edn {
namespace tc {

constexpr int kChunkA = 2048;            // bytes of one 8-wide K chunk of a 128-row bf16 operand tile

// ---- weight packing -----------------------------------------------------------------------------------------------
// dst element (n, k) of a [K][N] layer -> bf16 index (k/16)*(N*16) + ((k%16)/8)*(N*8) + n*8 + k%8
// col_rot: output column n reads source column (n + col_rot) % N (used to move the sigma column last).
static __global__ void pack_layer_kernel(const float* __restrict__ wt, int ld, int k_valid, int n_valid, int K, int N, int col_rot,
                                         __nv_bfloat16* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K * N) return;
  const int k = i / N, n = i - k * N;
  const int ns = (n + col_rot) % N;
  const float v = (k < k_valid && ns < n_valid) ? wt[(size_t)k * ld + ns] : 0.f;
  dst[(size_t)(k / 16) * (N * 16) + ((k % 16) / 8) * (N * 8) + n * 8 + (k % 8)] = __float2bfloat16_rn(v);
}

// ---- small device helpers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fast_sincos(float x, float* s, float* c) {
  // Cody-Waite reduction to [-pi, pi] then MUFU; abs error ~5e-7, far below bf16 resolution of the MMA operand
  const float k = rintf(x * 0.15915494309189535f);
  float r = fmaf(-k, 6.28125f, x);
  r = fmaf(-k, 1.9353071795864769e-3f, r);
  *s = __sinf(r);
  *c = __cosf(r);
}

// Packed fp32 arithmetic (sm_100 FFMA2 / FMUL2 / FADD2): two IEEE fp32 operations per instruction, same rounding as the scalar ones.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

// (x0, x1) fp32 -> packed bf16x2 hi = rn(x) and lo = rn(x - hi): x = hi + lo up to 2^-17 |x| (the operand split of the bf16 x 3 mode)
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(x0, x1);
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  lo = pack_bf16x2(x0 - h0, x1 - h1);
}

// Raw 16-byte channel chunk of a bf16 texel row / 2 x 16 bytes of an fp32 one.
template <typename T> struct Raw8;
template <> struct Raw8<__nv_bfloat16> {
  uint4 r;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { r = __ldg(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
};
template <> struct Raw8<float> {
  float4 lo, hi;
  __device__ __forceinline__ void load(const float* p) {
    lo = __ldg(reinterpret_cast<const float4*>(p)); hi = __ldg(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ void get(float (&v)[8]) const {
    v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
  }
};

// One gather task = 8 channels of one VM component at one point: 4 plane taps + 2 line taps (issue), then
// (bilinear plane) * (linear line) -> 16 bytes of bf16 (finish).  Issue and finish are split so that several tasks'
// loads are in flight together.
template <typename T>
struct GatherTask {
  Raw8<T> pv[4], lv[2];
  float pw[4], lw[2];
  __device__ __forceinline__ void issue(const T* __restrict__ plane, const T* __restrict__ line, int C, int c8,
                                        const Taps2& pt, const Taps1& lt) {
#pragma unroll
    for (int k = 0; k < 4; ++k) { pv[k].load(plane + (size_t)pt.off[k] * C + c8 * 8); pw[k] = pt.w[k]; }
#pragma unroll
    for (int k = 0; k < 2; ++k) { lv[k].load(line + (size_t)lt.off[k] * C + c8 * 8); lw[k] = lt.w[k]; }
  }
  __device__ __forceinline__ void finish(uint8_t* dst) const {
    float p[8], l[8], t[8];
    pv[0].get(t);
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = t[i] * pw[0];
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      pv[k].get(t);
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fmaf(t[i], pw[k], p[i]);
    }
    lv[0].get(t);
#pragma unroll
    for (int i = 0; i < 8; ++i) l[i] = t[i] * lw[0];
    lv[1].get(t);
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] *= fmaf(t[i], lw[1], l[i]);
    st_shared_v4(dst, pack_bf16x2(p[0], p[1]), pack_bf16x2(p[2], p[3]), pack_bf16x2(p[4], p[5]), pack_bf16x2(p[6], p[7]));
  }
  // fp32-grade variant (bf16 x 3 tensor-core parity mode): the fp32 product p is written as TWO bf16 operands, hi = bf16(p) and
  // lo = bf16(p - hi) (|p - hi - lo| <= 2^-17 |p|), 16 bytes each
  __device__ __forceinline__ void finish_split(uint8_t* dst_hi, uint8_t* dst_lo) const {
    float p[8], l[8], t[8];
    pv[0].get(t);
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = t[i] * pw[0];
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      pv[k].get(t);
#pragma unroll
      for (int i = 0; i < 8; ++i) p[i] = fmaf(t[i], pw[k], p[i]);
    }
    lv[0].get(t);
#pragma unroll
    for (int i = 0; i < 8; ++i) l[i] = t[i] * lw[0];
    lv[1].get(t);
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] *= fmaf(t[i], lw[1], l[i]);
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_bf16x2(p[2 * i], p[2 * i + 1], hi[i], lo[i]);
    st_shared_v4(dst_hi, hi[0], hi[1], hi[2], hi[3]);
    st_shared_v4(dst_lo, lo[0], lo[1], lo[2], lo[3]);
  }
  // Same arithmetic (fp32, one rounding per operation, identical results) on the packed fp32x2 pipe: 28 FMUL2 / FFMA2 instead of 56
  // scalar operations -- the gather warps' FMA-pipe time is what the epilogue warps compete with (fine_tc2.cu).
  __device__ __forceinline__ void finish2(uint8_t* dst) const {
    float2 p[4], l[4];
    float t[8];
    pv[0].get(t);
    const float2 w0 = make_float2(pw[0], pw[0]);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fmul2(make_float2(t[2 * i], t[2 * i + 1]), w0);
#pragma unroll
    for (int k = 1; k < 4; ++k) {
      pv[k].get(t);
      const float2 wk = make_float2(pw[k], pw[k]);
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = ffma2(make_float2(t[2 * i], t[2 * i + 1]), wk, p[i]);
    }
    lv[0].get(t);
    const float2 l0 = make_float2(lw[0], lw[0]), l1 = make_float2(lw[1], lw[1]);
#pragma unroll
    for (int i = 0; i < 4; ++i) l[i] = fmul2(make_float2(t[2 * i], t[2 * i + 1]), l0);
    lv[1].get(t);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = fmul2(p[i], ffma2(make_float2(t[2 * i], t[2 * i + 1]), l1, l[i]));
    st_shared_v4(dst, pack_bf16x2(p[0].x, p[0].y), pack_bf16x2(p[1].x, p[1].y), pack_bf16x2(p[2].x, p[2].y), pack_bf16x2(p[3].x, p[3].y));
  }
};


// Cooperative VM gather of ONE grid for point groups gi_begin .. gi_end - 1 (8 points each) of this warp's 32 points; `pos(pt, p)`
// returns the world position of tile row pt.  Results: (plane (.) line) products as bf16 into A chunks base .. base + 11 of row pt.
//   * Tap offsets / weights are computed ONCE per (point, component) -- lane 8 c + j: component c (0, 1, 2) of point 8 gi + j; lanes
//     24..31 idle -- and handed to the loading lanes with warp shuffles (36 SHFL instead of three ~70-instruction tap computations).
//   * Component 0 (64 channels = one 128-byte line per texel): lane (p4 = lane & 3, c8 = lane >> 2) reads chunk c8 of points
//     8 gi + p4 and 8 gi + 4 + p4, so every warp-wide load covers FOUR WHOLE lines (round 1: eight half lines; the L1 wavefront
//     count per byte bounds the gather).  Components 1 / 2 (16 channels): lane (q = lane >> 3, j = lane & 7) -> point 8 gi + j,
//     component 1 + q / 2, chunk q & 1.
template <typename T, typename Pos>
__device__ __forceinline__ void gather_points(const GridDev& g, uint8_t* As, const int base, const int gwarp, const int lane,
                                              const int gi_begin, const int gi_end, Pos pos) {
  const int p4 = lane & 3, c8 = lane >> 2, q = lane >> 3;
#pragma unroll 1
  for (int gi = gi_begin; gi < gi_end; ++gi) {
    Taps2 mp; Taps1 ml;
    {
      const int comp = min(lane >> 3, 2);
      float p[3], n[3];
      pos(gwarp * 32 + gi * 8 + (lane & 7), p);
      normalize_pt(g, p, n);
      // matMode = [[0,1],[0,2],[1,2]], vecMode = [2,1,0]   (voxnerf.py:99-100)
      const float px = comp == 2 ? n[1] : n[0], py = comp == 0 ? n[1] : n[2], lv = comp == 0 ? n[2] : (comp == 1 ? n[1] : n[0]);
      plane_taps(px, py, g.ph[comp], g.pw[comp], mp);
      line_taps(lv, g.ll[comp], ml);
    }
    auto fetch = [&](int src, Taps2& pt2, Taps1& lt1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { pt2.off[k] = __shfl_sync(0xffffffffu, mp.off[k], src); pt2.w[k] = __shfl_sync(0xffffffffu, mp.w[k], src); }
#pragma unroll
      for (int k = 0; k < 2; ++k) { lt1.off[k] = __shfl_sync(0xffffffffu, ml.off[k], src); lt1.w[k] = __shfl_sync(0xffffffffu, ml.w[k], src); }
    };
    {  // component 0: plane (x,y), line z
      GatherTask<T> t0, t1;
      const T* pl = reinterpret_cast<const T*>(g.plane[0]);
      const T* ln = reinterpret_cast<const T*>(g.line[0]);
      const int ptA = gwarp * 32 + gi * 8 + p4, ptB = ptA + 4;
      Taps2 pt2; Taps1 lt1;
      fetch(p4, pt2, lt1);
      t0.issue(pl, ln, 64, c8, pt2, lt1);
      fetch(p4 + 4, pt2, lt1);
      t1.issue(pl, ln, 64, c8, pt2, lt1);
      t0.finish2(As + ptA * 16 + (base + c8) * kChunkA);
      t1.finish2(As + ptB * 16 + (base + c8) * kChunkA);
    }
    {  // components 1 (plane (x,z), line y) and 2 (plane (y,z), line x): 16 channels each = 2 chunks each
      GatherTask<T> t2;
      const int pt = gwarp * 32 + gi * 8 + (lane & 7);
      const int comp = 1 + (q >> 1);
      Taps2 pt2; Taps1 lt1;
      fetch(comp * 8 + (lane & 7), pt2, lt1);
      t2.issue(reinterpret_cast<const T*>(g.plane[comp]), reinterpret_cast<const T*>(g.line[comp]), 16, q & 1, pt2, lt1);
      t2.finish2(As + pt * 16 + (base + 8 + q) * kChunkA);
    }
  }
}

// gather_points for the bf16 x 3 parity mode: fp32 (or bf16) grids, ONE task in flight per lane (an fp32 task holds 48 registers of
// raw taps), products written as hi / lo operand pairs into As_hi / As_lo.  Same lane mapping and shared tap computation.
template <typename T, typename Pos>
__device__ __forceinline__ void gather_points_split(const GridDev& g, uint8_t* As_hi, uint8_t* As_lo, const int base, const int gwarp,
                                                    const int lane, const int gi_begin, const int gi_end, Pos pos) {
  const int p4 = lane & 3, c8 = lane >> 2, q = lane >> 3;
#pragma unroll 1
  for (int gi = gi_begin; gi < gi_end; ++gi) {
    Taps2 mp; Taps1 ml;
    {
      const int comp = min(lane >> 3, 2);
      float p[3], n[3];
      pos(gwarp * 32 + gi * 8 + (lane & 7), p);
      normalize_pt(g, p, n);
      const float px = comp == 2 ? n[1] : n[0], py = comp == 0 ? n[1] : n[2], lv = comp == 0 ? n[2] : (comp == 1 ? n[1] : n[0]);
      plane_taps(px, py, g.ph[comp], g.pw[comp], mp);
      line_taps(lv, g.ll[comp], ml);
    }
    auto fetch = [&](int src, Taps2& pt2, Taps1& lt1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { pt2.off[k] = __shfl_sync(0xffffffffu, mp.off[k], src); pt2.w[k] = __shfl_sync(0xffffffffu, mp.w[k], src); }
#pragma unroll
      for (int k = 0; k < 2; ++k) { lt1.off[k] = __shfl_sync(0xffffffffu, ml.off[k], src); lt1.w[k] = __shfl_sync(0xffffffffu, ml.w[k], src); }
    };
    const T* pl = reinterpret_cast<const T*>(g.plane[0]);
    const T* ln = reinterpret_cast<const T*>(g.line[0]);
#pragma unroll 1
    for (int ab = 0; ab < 2; ++ab) {     // component 0: chunk c8 of points 8 gi + p4 and 8 gi + 4 + p4
      GatherTask<T> t;
      const int pt = gwarp * 32 + gi * 8 + p4 + 4 * ab;
      Taps2 pt2; Taps1 lt1;
      fetch(p4 + 4 * ab, pt2, lt1);
      t.issue(pl, ln, 64, c8, pt2, lt1);
      t.finish_split(As_hi + pt * 16 + (base + c8) * kChunkA, As_lo + pt * 16 + (base + c8) * kChunkA);
    }
    {
      GatherTask<T> t2;
      const int pt = gwarp * 32 + gi * 8 + (lane & 7);
      const int comp = 1 + (q >> 1);
      Taps2 pt2; Taps1 lt1;
      fetch(comp * 8 + (lane & 7), pt2, lt1);
      t2.issue(reinterpret_cast<const T*>(g.plane[comp]), reinterpret_cast<const T*>(g.line[comp]), 16, q & 1, pt2, lt1);
      t2.finish_split(As_hi + pt * 16 + (base + 8 + q) * kChunkA, As_lo + pt * 16 + (base + 8 + q) * kChunkA);
    }
  }
}

// One lane of a CONVERGED warp (all operands warp-uniform): lets the compiler keep descriptors in uniform registers and emit a
// plainly predicated UTCHMMA / UTCBAR / UBLKCP instead of the per-active-lane retry loop it wraps around them inside `if (lane == 0)`.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
  return pred != 0;
}

}  // namespace tc
}  // namespace edn