"""CPU: the C-ABI library loads and exports every symbol include/evdeblur_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from util import GOLDEN  # noqa: F401  (path setup)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "evdeblur_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(edn_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for s in ("edn_render_coarse_fwd", "edn_sample_pdf_merge", "edn_render_fine_fwd", "edn_pack_vm_plane"):
        assert s in syms


def test_library_exports_every_declared_symbol():
    from evdeblurnerf_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/evdeblur_b200.h but not exported"
    assert lib.edn_abi_version() == 1


def test_binding_table_matches_header():
    from evdeblurnerf_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_no_cpu_fallback():
    import torch
    from evdeblurnerf_b200 import RenderEngine
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError):
        RenderEngine({}, (0, 0, 0), (1, 1, 1))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "evdeblurnerf_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "evdeblur_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f
