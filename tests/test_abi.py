"""CPU: the C-ABI library builds, loads without a GPU and exports every symbol include/llamagen_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "llamagen_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lg_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_header_symbols():
    from llamagen_b200 import _lib
    lib = _lib.load()
    assert lib.lg_version() == 1
    names = _header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"


def test_no_libcuda_link_dependency():
    # the .so must be loadable on a CPU-only box: static cudart, no DT_NEEDED on libcuda
    from llamagen_b200 import _lib
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libcudart" not in out


def test_error_reporting_without_gpu():
    from llamagen_b200 import _lib
    lib = _lib.load()
    cfg = _lib.ModelCfg(2, 2, 100, 256, 512, 1, 16, 10, 64, _lib.LG_MODEL_C2I, _lib.LG_DTYPE_BF16, 1e-5)
    h = ctypes.c_void_p()
    rc = lib.lg_engine_create(ctypes.byref(cfg), 0, ctypes.byref(h))     # head_dim 50 is unsupported
    assert rc < 0 and b"head_dim" in lib.lg_last_error()
    cfg.dim = 128
    assert lib.lg_engine_create(ctypes.byref(cfg), 0, ctypes.byref(h)) == 0
    assert lib.lg_engine_finalize(h) < 0 and b"missing weight" in lib.lg_last_error()
    lib.lg_engine_destroy(h)


def test_product_never_imports_oracle():
    """The shipped package must not route through the CPU oracle (graded constraint ③)."""
    pkg = os.path.join(ROOT, "llamagen_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports oracle"
