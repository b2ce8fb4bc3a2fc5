"""ctypes loader for libdsurf_b200.so (the sm_100a CUDA implementation).

There is NO CPU fallback: if the shared library is missing, import fails; if no sm_100 GPU is
visible, every entry point returns DSURF_ERR_NO_CUDA, which is raised as :class:`DsurfError`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
# DSURF_B200_LIB: load another build of the same library (A/B runs of kernel variants); no fallback either way
LIB_PATH = os.environ.get("DSURF_B200_LIB") or os.path.join(HERE, "libdsurf_b200.so")
CSRC = os.path.join(HERE, "csrc")

OK = 0
ERR_NAMES = {
    1: "SOURCE_OUTSIDE", 2: "RECEIVER_OUTSIDE", 3: "NO_CUDA", 4: "CUDA", 5: "CAPACITY",
    6: "BAD_ARG", 7: "HEAP", 8: "NCCL",
}


class DsurfError(RuntimeError):
    def __init__(self, code: int, where: str, detail: str = ""):
        self.code = code
        super().__init__(f"{where}: DSURF_ERR_{ERR_NAMES.get(code, code)} {detail}")


def build(verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC, "-j", str(min(8, os.cpu_count() or 1))],
                         capture_output=not verbose, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libdsurf_b200.so failed:\n" + (out.stdout or "") + (out.stderr or ""))
    return LIB_PATH


_lib = None

f32p = C.POINTER(C.c_float)
f64p = C.POINTER(C.c_double)
i32p = C.POINTER(C.c_int)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C dsurftomo_b200/csrc` -- dsurftomo_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.dsurf_last_error.restype = C.c_char_p
    L.dsurf_build_info.restype = C.c_char_p
    L.dsurf_plan_nar.restype = C.c_int64
    L.dsurf_lsmr_nnz.restype = C.c_int64
    L.dsurf_plan_nar.argtypes = [C.c_void_p]
    L.dsurf_plan_last_sweeps_ms.restype = C.c_double
    L.dsurf_plan_last_sweeps_ms.argtypes = [C.c_void_p]
    L.dsurf_plan_last_gather_ms.restype = C.c_double
    L.dsurf_plan_last_gather_ms.argtypes = [C.c_void_p]
    L.dsurf_lsmr_nnz.argtypes = [C.c_void_p]
    for name in ("dsurf_plan_num_gathers", "dsurf_plan_nrows", "dsurf_plan_destroy", "dsurf_plan_dispersion",
                 "dsurf_plan_reset_rows", "dsurf_lsmr_destroy"):
        getattr(L, name).argtypes = [C.c_void_p]
    L.dsurf_plan_sweeps.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.dsurf_plan_set_raypath.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    L.dsurf_plan_raypath_count.argtypes = [C.c_void_p]
    L.dsurf_plan_raypath_count.restype = C.c_int64
    L.dsurf_plan_num_sweeps.argtypes = [C.c_void_p, C.c_int, C.c_int]
    _lib = L
    return L


def check(rc: int, where: str):
    if rc != OK:
        raise DsurfError(rc, where, lib().dsurf_last_error().decode(errors="replace"))


def ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


EXPORTED_SYMBOLS = [
    # gfortran-convention drop-ins
    "calsurfg_", "depthkernel_", "caldespersion_", "surfdisp96_", "__lsmrmodule_MOD_lsmr", "aprod_", "synthetic_",
    # neutral C API
    "dsurf_last_error", "dsurf_set_device", "dsurf_build_info", "dsurf_set_eikonal_mode", "dsurf_get_eikonal_mode", "dsurf_calsurfg", "dsurf_synthetic", "dsurf_depthkernel",
    "dsurf_surfdisp96", "dsurf_surfdisp96_batch", "dsurf_lsmr", "dsurf_aprod",
    "dsurf_plan_create", "dsurf_plan_create_forward", "dsurf_plan_destroy", "dsurf_plan_set_model", "dsurf_plan_dispersion",
    "dsurf_plan_set_map", "dsurf_plan_set_dispersion", "dsurf_plan_finalize_dispersion", "dsurf_plan_reset_rows", "dsurf_plan_sweeps", "dsurf_plan_num_gathers",
    "dsurf_plan_num_sweeps", "dsurf_plan_nar", "dsurf_plan_nrows", "dsurf_plan_download",
    "dsurf_plan_debug_sweep", "dsurf_plan_get_dispersion", "dsurf_plan_timings", "dsurf_plan_last_sweeps_ms",
    "dsurf_plan_glue_results", "dsurf_plan_update_model",
    "dsurf_lsmr_create", "dsurf_lsmr_create_from_plan", "dsurf_lsmr_hint_geometry", "dsurf_lsmr_destroy", "dsurf_lsmr_set_comm",
    "dsurf_lsmr_xchg_export", "dsurf_lsmr_xchg_attach",
    "dsurf_lsmr_solve", "dsurf_lsmr_nnz", "dsurf_nccl_unique_id", "dsurf_nccl_comm_init",
    "dsurf_nccl_comm_destroy",
]
