"""ctypes loader for oracle/_ref/libcluster_ref.so: the REFERENCE's own sources
(src/cluster.cpp, distributions.cpp, probutils.cpp, comutils.cpp) compiled against the
Eigen/Boost stand-ins in oracle/refshim.  TEST INFRASTRUCTURE ONLY (see oracle/README.md)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libcluster_ref.so")
_LIB = None


def available():
    return os.path.exists(SO)


def build():
    if os.path.isdir("/root/reference/src"):
        subprocess.call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return available()


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(SO)
        dp, lp, vp = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_void_p
        L.ref_last_error.restype = C.c_char_p
        L.ref_learn.argtypes = [C.c_int, C.c_int, dp, lp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_uint, C.POINTER(vp)]
        if hasattr(L, "ref_learn_wprior"):   # absent from a library built before round 2
            L.ref_learn_wprior.argtypes = [C.c_int, dp, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_int, C.c_uint, C.POINTER(vp)]
        L.ref_vbem.argtypes = [C.c_int, C.c_int, dp, lp, C.c_int, dp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.ref_free.argtypes = [vp]
        L.ref_F.restype = C.c_double
        L.ref_F.argtypes = [vp]
        for f in ("ref_K", "ref_J"):
            getattr(L, f).argtypes = [vp]
        L.ref_get_qZ.argtypes = [vp, C.c_int, dp]
        L.ref_qrows.restype = C.c_int64
        L.ref_qrows.argtypes = [vp, C.c_int]
        L.ref_qcols.argtypes = [vp, C.c_int]
        L.ref_get_weights.argtypes = [vp, C.c_int, dp, dp, dp]
        L.ref_weights_K.argtypes = [vp, C.c_int]
        L.ref_get_cluster.argtypes = [vp, C.c_int, dp, dp, dp, dp]
        L.ref_cov_len.argtypes = [vp, C.c_int]
        _LIB = L
    return _LIB


class RefError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("reference status %d: %s" % (code, msg))
        self.code = code


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Result:
    def __init__(self, h, D):
        L = lib()
        self.F = L.ref_F(h)
        self.K = L.ref_K(h)
        J = L.ref_J(h)
        self.qZ = []
        for j in range(J):
            q = np.zeros((L.ref_qrows(h, j), L.ref_qcols(h, j)))
            if q.size:
                L.ref_get_qZ(h, j, _dp(q))
            self.qZ.append(q)
        self.Elogweight, self.Nk, self.wfen = [], [], []
        for j in range(J):
            k = L.ref_weights_K(h, j)
            e, n, f = np.zeros(k), np.zeros(k), C.c_double()
            L.ref_get_weights(h, j, _dp(e), _dp(n), C.byref(f))
            self.Elogweight.append(e); self.Nk.append(n); self.wfen.append(f.value)
        self.means, self.covs, self.N, self.cfen = [], [], [], []
        for k in range(self.K):
            m, c = np.zeros(D), np.zeros(L.ref_cov_len(h, k))
            n, f = C.c_double(), C.c_double()
            L.ref_get_cluster(h, k, _dp(m), _dp(c), C.byref(n), C.byref(f))
            self.means.append(m); self.covs.append(c.reshape(D, D) if c.size == D * D and D > 1 else c)
            self.N.append(n.value); self.cfen.append(f.value)
        L.ref_free(h)


def _pack(groups):
    if isinstance(groups, np.ndarray):
        groups = [groups]
    groups = [np.ascontiguousarray(g, dtype=np.float64) for g in groups]
    Nj = np.array([g.shape[0] for g in groups], dtype=np.int64)
    cat = np.ascontiguousarray(np.concatenate(groups, 0))
    return cat, Nj, groups[0].shape[1], len(groups)


def learn(model, groups, prior=1.0, maxclusters=-1, sparse=False, nthreads=1, weight_prior=None):
    """learnXXX of the reference.  weight_prior (single-group models): the caller passes Dirichlet(alpha) /
    StickBreak(concentration) instead of a default-constructed weight object."""
    cat, Nj, D, J = _pack(groups)
    h = C.c_void_p()
    if weight_prior is not None:
        assert J == 1
        rc = lib().ref_learn_wprior(model, _dp(cat), int(Nj[0]), D, prior, weight_prior, maxclusters, nthreads, C.byref(h))
        if rc:
            msg = lib().ref_last_error().decode()
            lib().ref_free(h)
            raise RefError(rc, msg)
        return Result(h, D)
    rc = lib().ref_learn(model, J, _dp(cat), Nj.ctypes.data_as(C.POINTER(C.c_int64)), D, prior, maxclusters, int(sparse),
                         nthreads, C.byref(h))
    if rc:
        msg = lib().ref_last_error().decode()
        lib().ref_free(h)
        raise RefError(rc, msg)
    return Result(h, D)


def vbem(model, groups, q0, prior=1.0, maxit=-1, sparse=False, nthreads=1):
    cat, Nj, D, J = _pack(groups)
    q0 = np.ascontiguousarray(q0, dtype=np.float64)
    h = C.c_void_p()
    rc = lib().ref_vbem(model, J, _dp(cat), Nj.ctypes.data_as(C.POINTER(C.c_int64)), D, _dp(q0), q0.shape[1], prior, maxit,
                        int(sparse), int(nthreads), C.byref(h))
    if rc:
        msg = lib().ref_last_error().decode()
        lib().ref_free(h)
        raise RefError(rc, msg)
    return Result(h, D)
