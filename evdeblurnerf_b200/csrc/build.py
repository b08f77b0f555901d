"""Builds libevdeblur_b200.so in-tree with nvcc for sm_100a (no torch headers, no JIT cache).

    python evdeblurnerf_b200/csrc/build.py [--force]
"""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libevdeblur_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xptxas", "-v"]
FLAGS += os.environ.get("EDN_NVCC_EXTRA", "").split()        # dev builds, e.g. -DEDN_TC_ABLATE_BUILD=1


def sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    hdrs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(HERE, "..", "..", "include", "evdeblur_b200.h"))
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for s in sources():
        src, obj = os.path.join(HERE, s), os.path.join(objdir, s[:-3] + ".o")
        if force or _stale(obj, [src] + hdrs):
            jobs.append((src, obj))

    def cc(job):
        src, obj = job
        r = subprocess.run([NVCC, *FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
        return src, r

    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        for src, r in ex.map(cc, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}")
            with open(os.path.join(objdir, os.path.basename(src) + ".ptxas.log"), "w") as f:
                f.write(r.stderr)
    objs = [os.path.join(objdir, s[:-3] + ".o") for s in sources()]
    if force or jobs or _stale(LIB, objs):
        # cuBLAS carries the plain tall GEMMs of the backward pass (field_bwd.cu); resolved from the CUDA toolkit or, when
        # torch is already imported, from the libcublas.so.12 torch loaded
        r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                            "-L/usr/local/cuda/lib64", "-lcublas", "-lcublasLt", "-Xlinker", "-rpath=/usr/local/cuda/lib64"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
