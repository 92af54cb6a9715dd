"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/later_b200.h declares, and the C++ LATER.h wrappers carry the reference's mangled names.
No compute call is made here (there is no GPU in this container)."""
import ctypes as C
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _declared_symbols():
    text = (ROOT / "include" / "later_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(later_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = _declared_symbols()
    for must in ("later_b200_create", "later_b200_destroy", "later_b200_rgsqrf",
                 "later_b200_rgsqrf_host", "later_b200_rgsqrf_stream_in", "later_b200_ormqr", "later_b200_ormqr2",
                 "later_b200_panel_qr", "later_b200_tsqr_apply", "later_b200_workspace_bytes"):
        assert must in names


def test_library_exports_every_declared_symbol(built_lib):
    lib = C.CDLL(str(built_lib))
    for name in _declared_symbols():
        assert hasattr(lib, name), f"{name} declared in later_b200.h but not exported"


def test_python_binding_covers_the_header(built_lib):
    from later_b200 import _lib
    assert sorted(_lib.SYMBOLS) == _declared_symbols()


def test_cxx_wrappers_have_the_reference_mangled_names(built_lib):
    # names test/test_qr.cu of the reference leaves undefined
    out = subprocess.run(["nm", "-D", "--defined-only", str(built_lib)], capture_output=True,
                         text=True, check=True).stdout
    for sym in ("_Z12later_rgsqrf8cudaCtxtiiPfiS0_iS0_iP6__halfi",
                "_Z11later_ormqriiPfiS_iS_", "_Z12later_ormqr2iiPfiS_iS_",
                "_Z10startTimerv", "_Z9stopTimerv", "_Z21generateUniformMatrixPfii",
                "_Z5snormiiPf", "_Z9print_envv", "_Z6setEyeiiPfi", "_Z8clearTriciiPfi"):
        assert sym in out, sym


def test_create_without_gpu_fails_loudly(built_lib):
    import torch
    if torch.cuda.is_available():
        return
    lib = C.CDLL(str(built_lib))
    lib.later_b200_create.restype = C.c_int
    lib.later_b200_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
    h = C.c_void_p()
    rc = lib.later_b200_create(C.byref(h), 0, None)
    assert rc == -4 and not h.value     # LATER_B200_ENODEV: no CPU fallback exists


def test_sass_is_blackwell_native(built_lib):
    """The GEMM kernels really are tcgen05 + TMA (UTCHMMA / UTMALDG / LDTM in the SASS)."""
    obj = ROOT / "build" / "tc_gemm.o"
    sass = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
    assert "HMMA.16816" not in sass   # no legacy mma.sync path
