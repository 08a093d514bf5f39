"""ctypes binding of include/libcluster_b200.h (the C ABI of the engine).

The shared library is built in-tree by libcluster_b200/build.py for sm_100a.
There is no Python or CPU fallback: if the library is missing, loading fails
loudly; if no GPU is visible, lcb_create returns LCB_ECUDA.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LCB_LIB_PATH") or os.path.join(HERE, "_lib", "liblcb200.so")  # override: diagnostic builds

OK, EINVAL, ERUNTIME, EDOMAIN, ECUDA, ENOMEM = range(6)
VDP, BGMM, DGMM, GMC, SGMC, DGMC = range(6)
W_DIRICHLET, W_STICKBREAK, W_GDIRICHLET = range(3)
C_GAUSSWISH, C_NORMGAMMA = range(2)
ROW_MAJOR, COL_MAJOR = 0, 1
F32, F64 = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)
_vp = C.c_void_p
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, _dp, C.c_int64, C.c_void_p)

# name -> (restype, argtypes); every symbol include/libcluster_b200.h declares
SIGNATURES = {
    "lcb_last_error": (C.c_char_p, []),
    "lcb_version": (C.c_char_p, []),
    "lcb_device_count": (C.c_int, []),
    "lcb_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int]),
    "lcb_destroy": (None, [_vp]),
    "lcb_set_data": (C.c_int, [_vp, C.c_int, C.POINTER(_dp), _lp, C.c_int, _lp, C.c_int]),
    "lcb_set_data_device_f32": (C.c_int, [_vp, _vp, C.c_int64, C.c_int, C.c_int64, _vp, C.c_int]),
    "lcb_learn": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_uint, _dp, _ip]),
    "lcb_model_init": (C.c_int, [_vp, C.c_int, C.c_double, C.c_double, C.c_int]),
    "lcb_set_qz": (C.c_int, [_vp, _dp, C.c_int]),
    "lcb_set_labels_device": (C.c_int, [_vp, _vp, C.c_int]),
    "lcb_vbem": (C.c_int, [_vp, C.c_int, _dp, _ip]),
    "lcb_vbem_step": (C.c_int, [_vp, _dp]),
    "lcb_num_clusters": (C.c_int, [_vp]),
    "lcb_num_groups": (C.c_int, [_vp]),
    "lcb_num_rows": (C.c_int64, [_vp, C.c_int]),
    "lcb_get_qz": (C.c_int, [_vp, C.c_int, _dp, C.c_int64, C.c_int]),
    "lcb_get_group_weights": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp]),
    "lcb_get_cluster": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "lcb_trace_len": (C.c_int, [_vp]),
    "lcb_get_trace": (C.c_int, [_vp, _dp, _ip]),
    "lcb_get_step_timing": (C.c_int, [_vp, _dp]),
    "lcb_get_estep_detail": (C.c_int, [_vp, _dp]),
    "lcb_stream": (_vp, [_vp]),
    "lcb_selftest_host_packing": (C.c_int, []),
    "lcb_get_step_counts": (C.c_int, [_vp, _dp]),
    "lcb_selftest_stick_order": (C.c_int, [_dp, C.c_int, _ip]),
    "lcb_nccl_unique_id": (C.c_int, [C.c_char_p]),
    "lcb_comm_init_nccl": (C.c_int, [_vp, C.c_char_p, C.c_int, C.c_int]),
    "lcb_comm_init_host": (C.c_int, [_vp, ALLREDUCE_FN, _vp, C.c_int, C.c_int]),
    "lcb_weights_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_double]),
    "lcb_weights_destroy": (None, [_vp]),
    "lcb_weights_update": (C.c_int, [_vp, _dp, C.c_int]),
    "lcb_weights_size": (C.c_int, [_vp]),
    "lcb_weights_elogweight": (C.c_int, [_vp, _dp]),
    "lcb_weights_getnk": (C.c_int, [_vp, _dp]),
    "lcb_weights_fenergy": (C.c_double, [_vp]),
    "lcb_cluster_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_double, C.c_int]),
    "lcb_cluster_destroy": (None, [_vp]),
    "lcb_cluster_addobs": (C.c_int, [_vp, _vp, _dp, _dp, C.c_int64, C.c_int64, C.c_int]),
    "lcb_cluster_update": (C.c_int, [_vp]),
    "lcb_cluster_clearobs": (C.c_int, [_vp]),
    "lcb_cluster_eloglike": (C.c_int, [_vp, _vp, _dp, C.c_int64, C.c_int64, C.c_int, _dp]),
    "lcb_cluster_splitobs": (C.c_int, [_vp, _vp, _dp, C.c_int64, C.c_int64, C.c_int, C.POINTER(C.c_uint8)]),
    "lcb_cluster_fenergy": (C.c_double, [_vp]),
    "lcb_cluster_getn": (C.c_double, [_vp]),
    "lcb_cluster_getprior": (C.c_double, [_vp]),
    "lcb_cluster_dim": (C.c_int, [_vp]),
    "lcb_cluster_getmean": (C.c_int, [_vp, _dp]),
    "lcb_cluster_getcov": (C.c_int, [_vp, _dp]),
    "lcb_cluster_get_stats": (C.c_int, [_vp, _dp, _dp, _dp]),
    "lcb_cluster_set_stats": (C.c_int, [_vp, C.c_double, _dp, _dp]),
    "lcb_packed_len": (C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "lcb_host_mstep": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp]),
    "lcb_shard_rows": (None, [C.c_int64, C.c_int, C.c_int, _lp, _lp]),
}

_LIB = None


class InvalidArgument(ValueError):
    """std::invalid_argument of the reference (LCB_EINVAL)."""


class FreeEnergyError(RuntimeError):
    """std::runtime_error of the reference, e.g. 'Free energy increase!' (LCB_ERUNTIME)."""


class DomainError(ArithmeticError):
    """std::domain_error out of probutils::logdet for a non-PD matrix (LCB_EDOMAIN)."""


class CudaError(RuntimeError):
    """Device, driver or kernel-image failure; there is no CPU fallback (LCB_ECUDA)."""


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libcluster_b200 native library not built: %s is missing. Run "
                "`python -m libcluster_b200.build` (or __graft_entry__.build()). "
                "There is no Python/CPU fallback." % LIB_PATH)
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            f = getattr(L, name)  # AttributeError if the ABI lost a symbol
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


def check(rc):
    if rc == OK:
        return
    msg = lib().lcb_last_error().decode("utf-8", "replace")
    if rc == EINVAL:
        raise InvalidArgument(msg)
    if rc == ERUNTIME:
        raise FreeEnergyError(msg)
    if rc == EDOMAIN:
        raise DomainError(msg)
    if rc == ENOMEM:
        raise MemoryError(msg)
    raise CudaError(msg)
