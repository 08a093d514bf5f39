"""ctypes loader for the CPU oracle (oracle/vb_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by the
product package libcluster_b200.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

VDP, BGMM, DGMM, GMC, SGMC, DGMC = range(6)
W_DIRICHLET, W_STICKBREAK, W_GDIRICHLET = range(3)
C_GAUSSWISH, C_NORMGAMMA = range(2)
MODEL_KINDS = {
    VDP: (W_STICKBREAK, C_GAUSSWISH),
    BGMM: (W_DIRICHLET, C_GAUSSWISH),
    DGMM: (W_DIRICHLET, C_NORMGAMMA),
    GMC: (W_GDIRICHLET, C_GAUSSWISH),
    SGMC: (W_DIRICHLET, C_GAUSSWISH),
    DGMC: (W_GDIRICHLET, C_NORMGAMMA),
}


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("oracle status %d: %s" % (code, msg))
        self.code = code


def build(force=False):
    so = os.path.join(_HERE, "libvb_oracle.so")
    src = os.path.join(_HERE, "vb_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libvb_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp = C.POINTER(C.c_double)
        ip = C.POINTER(C.c_int)
        lp = C.POINTER(C.c_int64)
        vp = C.c_void_p
        L.orc_last_error.restype = C.c_char_p
        L.orc_digamma.restype = C.c_double
        L.orc_digamma.argtypes = [C.c_double]
        L.orc_model_create.restype = vp
        L.orc_model_create.argtypes = [C.c_int, C.c_int, dp, lp, C.c_int]
        L.orc_model_destroy.argtypes = [vp]
        L.orc_learn.argtypes = [vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_uint]
        L.orc_vbem.argtypes = [vp, dp, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int]
        L.orc_vbem_iteration.argtypes = [vp, C.c_double, dp]
        L.orc_F.restype = C.c_double
        L.orc_F.argtypes = [vp]
        for f in ("orc_K", "orc_qK", "orc_trace_len"):
            getattr(L, f).argtypes = [vp]
        L.orc_trace.argtypes = [vp, dp, ip]
        L.orc_get_qZ.argtypes = [vp, dp]
        L.orc_get_weights.argtypes = [vp, C.c_int, dp, dp]
        L.orc_weights_K.argtypes = [vp, C.c_int]
        L.orc_weights_fenergy.restype = C.c_double
        L.orc_weights_fenergy.argtypes = [vp, C.c_int]
        L.orc_get_cluster.argtypes = [vp, C.c_int, dp, dp, dp, dp, dp, dp]
        L.orc_cluster_fenergy.restype = C.c_double
        L.orc_cluster_fenergy.argtypes = [vp, C.c_int]
        # operator handles
        L.orc_weight_new.restype = vp
        L.orc_weight_new.argtypes = [C.c_int, C.c_double]
        L.orc_weight_del.argtypes = [vp]
        L.orc_weight_update.argtypes = [vp, dp, C.c_int]
        L.orc_weight_K.argtypes = [vp]
        L.orc_weight_elogweight.argtypes = [vp, dp]
        L.orc_weight_getNk.argtypes = [vp, dp]
        L.orc_weight_fenergy.restype = C.c_double
        L.orc_weight_fenergy.argtypes = [vp]
        L.orc_cluster_new.restype = vp
        L.orc_cluster_new.argtypes = [C.c_int, C.c_double, C.c_int]
        L.orc_cluster_del.argtypes = [vp]
        L.orc_cluster_addobs.argtypes = [vp, dp, dp, C.c_int64]
        L.orc_cluster_update.argtypes = [vp]
        L.orc_cluster_clearobs.argtypes = [vp]
        L.orc_cluster_eloglike.argtypes = [vp, dp, C.c_int64, dp]
        L.orc_cluster_fenergy1.restype = C.c_double
        L.orc_cluster_fenergy1.argtypes = [vp]
        L.orc_cluster_splitobs.argtypes = [vp, dp, C.c_int64, C.POINTER(C.c_uint8)]
        L.orc_cluster_getN.restype = C.c_double
        L.orc_cluster_getN.argtypes = [vp]
        L.orc_cluster_get.argtypes = [vp, dp, dp, dp, dp, dp, dp]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _check(rc):
    if rc != 0:
        raise OracleError(rc, lib().orc_last_error().decode())


def digamma(x):
    return lib().orc_digamma(float(x))


class Model:
    """One fit of the restated cluster<W,C>() / vbem<W,C>() on groups of rows."""

    def __init__(self, model, groups):
        if isinstance(groups, np.ndarray):
            groups = [groups]
        self.model = model
        self.groups = [np.ascontiguousarray(g, dtype=np.float64) for g in groups]
        self.D = self.groups[0].shape[1]
        self.Nj = np.array([g.shape[0] for g in self.groups], dtype=np.int64)
        self.N = int(self.Nj.sum())
        cat = np.ascontiguousarray(np.concatenate(self.groups, axis=0))
        self._h = lib().orc_model_create(model, len(self.groups), _dp(cat),
                                        self.Nj.ctypes.data_as(C.POINTER(C.c_int64)), self.D)
        self.ckind = MODEL_KINDS[model][1]

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_model_destroy(self._h)
            self._h = None

    def learn(self, prior=1.0, weight_prior=-1.0, maxclusters=-1, sparse=False, nthreads=1):
        _check(lib().orc_learn(self._h, prior, weight_prior, maxclusters, int(sparse), nthreads))
        return self.F

    def vbem(self, q0, prior=1.0, weight_prior=-1.0, maxit=-1, sparse=False):
        q0 = np.ascontiguousarray(q0, dtype=np.float64)
        assert q0.shape[0] == self.N
        _check(lib().orc_vbem(self._h, _dp(q0), q0.shape[1], prior, weight_prior, maxit, int(sparse)))
        return self.F

    def iteration(self, prior=1.0):
        F = C.c_double()
        _check(lib().orc_vbem_iteration(self._h, prior, C.byref(F)))
        return F.value

    @property
    def F(self):
        return lib().orc_F(self._h)

    @property
    def K(self):
        return lib().orc_K(self._h)

    def trace(self):
        n = lib().orc_trace_len(self._h)
        F = np.zeros(n)
        K = np.zeros(n, dtype=np.int32)
        if n:
            lib().orc_trace(self._h, _dp(F), K.ctypes.data_as(C.POINTER(C.c_int)))
        return F, K

    def qZ(self):
        K = lib().orc_qK(self._h)
        out = np.zeros((self.N, K))
        lib().orc_get_qZ(self._h, _dp(out))
        return out

    def weights(self, j=0):
        K = lib().orc_weights_K(self._h, j)
        e = np.zeros(K)
        n = np.zeros(K)
        lib().orc_get_weights(self._h, j, _dp(e), _dp(n))
        return e, n

    def weights_fenergy(self, j=0):
        return lib().orc_weights_fenergy(self._h, j)

    def cluster(self, k):
        D = self.D
        S = D * D if self.ckind == C_GAUSSWISH else D
        scal = np.zeros(4)
        mean = np.zeros(D)
        iW = np.zeros(S)
        Ns = C.c_double()
        xs = np.zeros(D)
        xxs = np.zeros(S)
        lib().orc_get_cluster(self._h, k, _dp(scal), _dp(mean), _dp(iW), C.byref(Ns), _dp(xs), _dp(xxs))
        shp = (D, D) if self.ckind == C_GAUSSWISH else (D,)
        return dict(N=scal[0], nu=scal[1], beta=scal[2], logdW=scal[3], m=mean,
                    iW=iW.reshape(shp), N_s=Ns.value, x_s=xs, xx_s=xxs.reshape(shp),
                    fenergy=lib().orc_cluster_fenergy(self._h, k))


class Weight:
    def __init__(self, kind, prior=-1.0):
        self._h = lib().orc_weight_new(kind, prior)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_weight_del(self._h)
            self._h = None

    def update(self, Nk):
        Nk = np.ascontiguousarray(Nk, dtype=np.float64)
        lib().orc_weight_update(self._h, _dp(Nk), Nk.size)

    def Elogweight(self):
        out = np.zeros(lib().orc_weight_K(self._h))
        lib().orc_weight_elogweight(self._h, _dp(out))
        return out

    def getNk(self):
        out = np.zeros(lib().orc_weight_K(self._h))
        lib().orc_weight_getNk(self._h, _dp(out))
        return out

    def fenergy(self):
        return lib().orc_weight_fenergy(self._h)


class Cluster:
    def __init__(self, kind, prior, D):
        self.kind, self.D = kind, D
        self._h = lib().orc_cluster_new(kind, prior, D)
        if not self._h:
            raise OracleError(1, lib().orc_last_error().decode())

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_cluster_del(self._h)
            self._h = None

    def addobs(self, qk, X):
        qk = np.ascontiguousarray(qk, dtype=np.float64)
        X = np.ascontiguousarray(X, dtype=np.float64)
        lib().orc_cluster_addobs(self._h, _dp(qk), _dp(X), X.shape[0])

    def update(self):
        _check(lib().orc_cluster_update(self._h))

    def clearobs(self):
        lib().orc_cluster_clearobs(self._h)

    def Eloglike(self, X):
        X = np.ascontiguousarray(X, dtype=np.float64)
        out = np.zeros(X.shape[0])
        _check(lib().orc_cluster_eloglike(self._h, _dp(X), X.shape[0], _dp(out)))
        return out

    def fenergy(self):
        return lib().orc_cluster_fenergy1(self._h)

    def splitobs(self, X):
        X = np.ascontiguousarray(X, dtype=np.float64)
        out = np.zeros(X.shape[0], dtype=np.uint8)
        lib().orc_cluster_splitobs(self._h, _dp(X), X.shape[0], out.ctypes.data_as(C.POINTER(C.c_uint8)))
        return out.astype(bool)

    def getN(self):
        return lib().orc_cluster_getN(self._h)

    def state(self):
        D = self.D
        S = D * D if self.kind == C_GAUSSWISH else D
        scal = np.zeros(4)
        mean = np.zeros(D)
        iW = np.zeros(S)
        Ns = C.c_double()
        xs = np.zeros(D)
        xxs = np.zeros(S)
        lib().orc_cluster_get(self._h, _dp(scal), _dp(mean), _dp(iW), C.byref(Ns), _dp(xs), _dp(xxs))
        shp = (D, D) if self.kind == C_GAUSSWISH else (D,)
        return dict(N=scal[0], nu=scal[1], beta=scal[2], logdW=scal[3], m=mean,
                    iW=iW.reshape(shp), N_s=Ns.value, x_s=xs, xx_s=xxs.reshape(shp))
