"""ctypes binding of libdpiso.so (the C ABI declared in include/dpiso.h).

The product path has NO fallback: if the library cannot be loaded (or built with nvcc) importing this module raises,
and every call raises `DpisoError` on a non-zero status.  Tensors are passed as raw device pointers together with the
current torch CUDA stream; PyTorch is only the allocator / stream owner here.
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# DPISO_LIBRARY: load another build of the same C ABI (the diagnostics build of build.py --timing); no effect on the
# no-fallback rule: whatever is loaded must export every entry point below
_LIB_PATH = os.environ.get("DPISO_LIBRARY") or os.path.join(_HERE, "libdpiso.so")


class DpisoError(RuntimeError):
    pass


def _load():
    if not os.path.exists(_LIB_PATH):
        import importlib.util
        spec = importlib.util.spec_from_file_location("dpiso_build", os.path.join(_HERE, "..", "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    return C.CDLL(_LIB_PATH)


lib = _load()


class BicgTables(C.Structure):
    _fields_ = [("n", C.c_int), ("n_levels", C.c_int), ("wa", C.c_int), ("max_level", C.c_int),
                ("wl", C.c_int), ("wu", C.c_int), ("dx", C.c_int), ("rows_ok", C.c_int), ("level_ptr", C.c_void_p),
                ("perm", C.c_void_p), ("a_col", C.c_void_p), ("a_src", C.c_void_p), ("a_rev", C.c_void_p),
                ("r_col", C.c_void_p), ("r_src", C.c_void_p), ("r_rev", C.c_void_p), ("c_lsrc", C.c_void_p),
                ("c_lrev", C.c_void_p), ("c_usrc", C.c_void_p), ("c_lfar", C.c_void_p), ("c_ufar", C.c_void_p),
                ("c_dsrc", C.c_void_p), ("m_nbr", C.c_void_p), ("m_lfar", C.c_void_p), ("m_ufar", C.c_void_p),
                ("owner", C.c_void_p), ("owner_is_host", C.c_int), ("sym", C.c_int), ("band_ok", C.c_int),
                ("far", C.c_int * 8)]


_I, _F, _P, _SZ = C.c_int, C.c_float, C.c_void_p, C.c_size_t
_SIGS = {
    "dpiso_version": ([], _I),
    "dpiso_last_error": ([], C.c_char_p),
    "dpiso_sizes": ([_I, _I, _I, _I, _P, _P], _I),
    "dpiso_csr_structure": ([_I, _I, _I, _I, _P, _P, _P], _I),
    "dpiso_assemble": ([_I, _I, _I, _I, _I, _F, _F, _F, _F, _F, _P, _P, _P, _P, _P, _I, _P, _P, _P], _I),
    "dpiso_predictor_rhs": ([_I, _I, _I, _F, _F, _F, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P], _I),
    "dpiso_fv_gradient": ([_I, _I, _I, _F, _F, _P, _P, _P, _P, _P], _I),
    "dpiso_fv_divergence": ([_I, _I, _I, _F, _F, _P, _P, _F, _P, _P], _I),
    "dpiso_corrector1": ([_I, _I, _I, _F, _F, _F, _P, _P, _P, _P, _P, _P, _P], _I),
    "dpiso_h_apply": ([_I, _I, _I, _I, _I, _F, _P, _P, _P, _P, _P, _P], _I),
    "dpiso_corrector2": ([_I, _I, _I, _F, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P], _I),
    "dpiso_fv_gradient_adj": ([_I, _I, _I, _F, _F, _P, _P, _P, _P, _F, _F, _I, _P, _P, _P], _I),
    "dpiso_fv_divergence_adj": ([_I, _I, _I, _I, _I, _F, _F, _P, _P, _P, _P, _F, _P, _P], _I),
    "dpiso_h_apply_adj": ([_I, _P, _P, _I, _I, _F, _P, _P, _P, _P, _P, _P, _P], _I),
    "dpiso_predictor_rhs_adj": ([_I, _I, _I, _F, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P], _I),
    "dpiso_bicg_tables_create": ([_I, _I, _I, _I, _I, _I, _P, _P], _I),
    "dpiso_bicg_tables_create_host": ([_I, _I, _I, _I, _I, _I, _P], _I),
    "dpiso_bicg_tables_destroy": ([_P], _I),
    "dpiso_bicgstab_workspace_floats": ([_P, _P], _SZ),
    "dpiso_bicgstab_set_timing": ([_P], _I),
    "dpiso_bicgstab_set_debug": ([_I], _I),
    "dpiso_bicgstab_set_reuse_policy": ([_I], _I),
    "dpiso_bicgstab_set_band_cluster": ([_I], _I),
    "dpiso_bicgstab_set_tile_cluster": ([_I], _I),
    "dpiso_bicgstab_supports_factor_reuse": ([_P, _P], _I),
    "dpiso_bicgstab_f64_workspace_bytes": ([_P, _P], _SZ),
    "dpiso_bicgstab_ilu_f64": ([_I, _P, _P, _I, _I, _P, _I, _P, _P, _F, _I, _P, _P, _P, _P, _P], _I),
    "dpiso_bicgstab_ilu": ([_I, _P, _P, _I, _I, _P, _I, _P, _P, _F, _I, _P, _P, _P, _P, _P, _P, _P], _I),
    "dpiso_laplace_f64": ([_I, _I, _I, _P, _P, _P, _I, _F, _F, _P, _P], _I),
    "dpiso_laplace_f32": ([_I, _I, _I, _P, _P, _P, _I, _F, _F, _P, _P], _I),
    "dpiso_pressure_cg_workspace_bytes": ([_I, _I, _I, _I, _I], _SZ),
    "dpiso_pressure_cg_f64": ([_I, _I, _I, _I, _I, _P, _P, _F, _I, _I, _I, _P, _P, _P, _P, _P], _I),
    "dpiso_pressure_cg_f32": ([_I, _I, _I, _I, _I, _P, _P, _F, _I, _I, _I, _P, _P, _P, _P, _P], _I),
    "dpiso_pressure_cg_mixed": ([_I, _I, _I, _I, _I, _P, _P, _F, _I, _I, _I, _P, _P, _P, _P], _I),
    "dpiso_pressure_cg_last_config": ([_P], _I),
    "dpiso_pressure_cg_set_tuning": ([_I, _I], _I),
    "dpiso_pressure_cg_set_reduction_order": ([_I], _I),
    "dpiso_pressure_cg_set_static_nx": ([_I], _I),
}
for _name, (_args, _res) in _SIGS.items():
    _fn = getattr(lib, _name)
    _fn.argtypes = _args
    _fn.restype = _res

EXPORTS = tuple(_SIGS)


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise DpisoError("libdpiso needs CUDA tensors (got a %s tensor): there is no CPU path" % t.device)
    if not t.is_contiguous():
        raise DpisoError("libdpiso needs contiguous tensors")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def check(rc, what):
    if rc != 0:
        raise DpisoError("%s failed (%d): %s" % (what, rc, lib.dpiso_last_error().decode()))


def int4(values):
    return (C.c_int * 4)(*[int(v) for v in values])
