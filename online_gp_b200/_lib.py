"""ctypes binding of libwiski_b200.so (the C ABI declared in include/wiski_b200.h).

There is no CPU fallback anywhere in this package: if the shared library is missing, or a tensor is not on a CUDA
device, the call fails loudly.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libwiski_b200.so")

_lib = None

_P = c_void_p          # device pointer
_I64P = POINTER(c_int64)
_S = c_void_p          # cudaStream_t


def _sigs(real, realp):
    """(name, argtypes) for one dtype; `real` = c_float/c_double, `realp` = host pointer to it."""
    return {
        "wiski_interp_fwd": [_P, c_int64, c_int, _I64P, realp, realp, realp, realp, realp, realp, _P, _P, _P, _S],
        "wiski_interp_bwd": [_P, c_int64, c_int, _I64P, realp, realp, realp, realp, _P, _P, _S],
        "wiski_gather": [_P, _P, c_int64, c_int64, _P, c_int64, c_int64, _P, _S],
        "wiski_scatter_add": [_P, _P, c_int64, c_int64, _P, c_int64, c_int64, _P, _S],
        "wiski_kron_toeplitz_mm": [_P, c_int, _I64P, c_int64, _P, c_int64, _P, _P, _S],
        "wiski_kron_toeplitz_bwd_cols": [_P, c_int, _I64P, c_int64, _P, _P, c_int64, _P, _P, _S],
        "wiski_kron_axis_apply": [_P, _P, _P, c_int64, c_int64, c_int64, _S],
        "wiski_kron_axis_contract": [_P, _P, c_int64, c_int64, c_int64, _P, _S],
        "wiski_panel_rmul": [_P, c_int64, c_int64, _P, c_int64, _P, _S],
        "wiski_panel_lowrank_update": [_P, c_int64, c_int64, _P, _P, c_int64, _S],
        "wiski_panel_lowrank_update2": [_P, _P, c_int64, c_int64, _P, _P, _P, c_int64, _S],
        "wiski_panel_outer_add": [_P, c_int64, c_int64, _P, c_int64, _P, _S],
        "wiski_panel_lowrank_update2_t": [_P, _P, c_int64, c_int64, _P, _P, _P, c_int64, _P, _S],
        "wiski_gram": [_P, _P, c_int64, c_int64, c_int64, _P, _P, _S],
        "wiski_q_matvec": [_P, _P, c_int64, c_int64, _P, c_int64, _P, _P, _S],
        "wiski_cg_solve": [_P, _P, c_int64, c_int64, _P, c_int64, real, c_int, c_int, _P, POINTER(c_int), realp, _P, _S],
    }


#: every symbol include/wiski_b200.h declares (tests check the .so exports all of them)
EXPORTED = ["wiski_last_error", "wiski_abi_version", "wiski_launch_count", "wiski_kron_fused_supported",
            "wiski_kron_fused_pair_apply_f32", "wiski_kron_fused_pair_grad_f32", "wiski_kron_fused_pair_grad_dir_f32", "wiski_gram_work_elems", "wiski_qmv_work_elems",
            "wiski_cg_work_elems", "wiski_kron_toeplitz_bwd_work_elems", "wiski_kron_axis_tc_work_elems",
            "wiski_kron_axis_apply_tc_f32", "wiski_kron_axis_contract_tc_f32", "wiski_kron_fused_pair_apply_lay_f32",
            "wiski_kron_fused_pair_grad_lay_f32", "wiski_gram_chunked_f32", "wiski_panel_rmul_chunked_f32",
            "wiski_kron_fused_pair_grad_dir_lay_f32", "wiski_kron_tc_enable", "wiski_gram_sym_f32",
            "wiski_panel_rmul_ex_f32", "wiski_gram_chunked_sym_f32",
            "wiski_kron_pair_apply_axes_f32", "wiski_kron_pair_grad_dir_axes_f32", "wiski_panel_rmul_ex_work_elems",
            "wiski_set_background", "wiski_kron_pair_apply_push_f32", "wiski_kron_pair_grad_dir_push_f32", "wiski_panel_rmul_push_f32"] + [
    f"{n}_{sfx}" for n in _sigs(c_float, POINTER(c_float)) for sfx in ("f32", "f64")]


def load():
    """Load the shared library once; raise if it has not been built (python __graft_entry__.py / csrc/build.sh)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"online_gp_b200: CUDA library not built: {LIB_PATH} is missing. Run online_gp_b200/csrc/build.sh "
            f"(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    lib.wiski_last_error.restype = c_char_p
    lib.wiski_last_error.argtypes = []
    lib.wiski_abi_version.restype = c_int
    lib.wiski_launch_count.restype = ctypes.c_longlong
    lib.wiski_launch_count.argtypes = []
    for name in ("wiski_gram_work_elems", "wiski_qmv_work_elems", "wiski_cg_work_elems"):
        fn = getattr(lib, name)
        fn.restype = c_int64
        fn.argtypes = [c_int64, c_int64, c_int64]
    lib.wiski_kron_toeplitz_bwd_work_elems.restype = c_int64
    lib.wiski_kron_toeplitz_bwd_work_elems.argtypes = [c_int, c_int64, c_int64, c_int64, c_int]
    lib.wiski_kron_fused_supported.restype = c_int
    lib.wiski_kron_fused_supported.argtypes = [c_int, _I64P, c_int64]
    lib.wiski_kron_fused_pair_apply_f32.restype = c_int
    lib.wiski_kron_fused_pair_apply_f32.argtypes = [_P, c_int, _I64P, c_int64, c_int, _P, _P, c_int64, _S]
    lib.wiski_kron_fused_pair_grad_f32.restype = c_int
    lib.wiski_kron_fused_pair_grad_f32.argtypes = [_P, c_int, _I64P, c_int64, c_int, _P, _P, _P, c_int64, _P, _P, _S]
    lib.wiski_kron_fused_pair_grad_dir_f32.restype = c_int
    lib.wiski_kron_fused_pair_grad_dir_f32.argtypes = [_P, _P, c_int, _I64P, c_int64, c_int, _P, _P, _P, c_int64, _P, _S]
    lib.wiski_kron_fused_pair_grad_dir_lay_f32.restype = c_int
    lib.wiski_kron_fused_pair_grad_dir_lay_f32.argtypes = [_P, _P, c_int, _I64P, c_int64, c_int, _P, _P, _P, c_int64, _P, _I64P, _S]
    lib.wiski_kron_pair_apply_axes_f32.restype = c_int
    lib.wiski_kron_pair_apply_axes_f32.argtypes = [_P, c_int, _I64P, c_int64, c_int, c_int, _P, _P, c_int64, _I64P, _S]
    lib.wiski_kron_pair_grad_dir_axes_f32.restype = c_int
    lib.wiski_kron_pair_grad_dir_axes_f32.argtypes = [_P, _P, c_int, _I64P, c_int64, c_int, c_int, _P, _P, _P, c_int64, _P, _I64P, _S]
    _PP = POINTER(c_void_p)
    lib.wiski_kron_pair_apply_push_f32.restype = c_int
    lib.wiski_kron_pair_apply_push_f32.argtypes = [_P, c_int, _I64P, c_int64, c_int, c_int, _P, c_int64, _I64P, _PP, c_int, c_int, _S]
    lib.wiski_kron_pair_grad_dir_push_f32.restype = c_int
    lib.wiski_kron_pair_grad_dir_push_f32.argtypes = [_P, _P, c_int, _I64P, c_int64, c_int, c_int, _P, _P, c_int64, _P, _I64P, _PP,
                                                      c_int, _S]
    lib.wiski_panel_rmul_push_f32.restype = c_int
    lib.wiski_panel_rmul_push_f32.argtypes = [_P, c_int64, c_int64, _P, c_int64, c_int, _PP, c_int, _P, _S]
    lib.wiski_gram_chunked_sym_f32.restype = c_int
    lib.wiski_gram_chunked_sym_f32.argtypes = [_P, _P, c_int64, c_int64, c_int64, _P, _P, _S]
    lib.wiski_gram_sym_f32.restype = c_int
    lib.wiski_gram_sym_f32.argtypes = [_P, _P, c_int64, c_int64, _P, _P, _S]
    lib.wiski_panel_rmul_ex_f32.restype = c_int
    lib.wiski_panel_rmul_ex_f32.argtypes = [_P, c_int64, c_int64, _P, c_int64, c_int64, c_int, _P, _P, _S]
    lib.wiski_panel_rmul_ex_work_elems.restype = c_int64
    lib.wiski_panel_rmul_ex_work_elems.argtypes = [c_int64, c_int64]
    lib.wiski_set_background.restype = c_int
    lib.wiski_set_background.argtypes = [c_int]
    lib.wiski_kron_tc_enable.restype = c_int
    lib.wiski_kron_tc_enable.argtypes = [c_int]
    lib.wiski_kron_fused_pair_apply_lay_f32.restype = c_int
    lib.wiski_kron_fused_pair_apply_lay_f32.argtypes = [_P, c_int, _I64P, c_int64, c_int, _P, _P, c_int64, _I64P, _S]
    lib.wiski_kron_fused_pair_grad_lay_f32.restype = c_int
    lib.wiski_kron_fused_pair_grad_lay_f32.argtypes = [_P, c_int, _I64P, c_int64, c_int, _P, _P, _P, c_int64, _P, _P, _I64P, _S]
    lib.wiski_gram_chunked_f32.restype = c_int
    lib.wiski_gram_chunked_f32.argtypes = [_P, _P, c_int64, c_int64, c_int64, c_int64, _P, _P, _S]
    lib.wiski_panel_rmul_chunked_f32.restype = c_int
    lib.wiski_panel_rmul_chunked_f32.argtypes = [_P, c_int64, c_int64, _P, c_int64, c_int64, _P, _S]
    lib.wiski_kron_axis_tc_work_elems.restype = c_int64
    lib.wiski_kron_axis_tc_work_elems.argtypes = [c_int64, c_int64, c_int64, c_int]
    lib.wiski_kron_axis_apply_tc_f32.restype = c_int
    lib.wiski_kron_axis_apply_tc_f32.argtypes = [_P, _P, _P, c_int64, c_int64, c_int64, _P, _S]
    lib.wiski_kron_axis_contract_tc_f32.restype = c_int
    lib.wiski_kron_axis_contract_tc_f32.argtypes = [_P, _P, c_int64, c_int64, c_int64, _P, _P, _S]
    for sfx, real in (("f32", c_float), ("f64", c_double)):
        for name, args in _sigs(real, POINTER(real)).items():
            fn = getattr(lib, f"{name}_{sfx}")
            fn.restype = c_int
            fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().wiski_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"online_gp_b200.{what} failed (status {rc}): {msg}")
