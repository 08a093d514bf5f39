"""Python face of the engine: mirrors the reference's own Python module
(python/libclusterpy.cpp:135-241) and its C++ operator classes
(include/distributions.h) on top of the C ABI.

    f, qZ, w, mu, cov = learnBGMM(X, prior=1.0, maxclusters=-1)

Same names, argument order and return tuples as libclusterpy.learnVDP/BGMM/
GMC/SGMC (python/libclusterpy.h:306-381), plus learnDGMM/DGMC which the
reference exposes only from C++ (include/libcluster.h:262,462).
"""
import ctypes as C

import numpy as np

from . import _native as nat
from ._native import (BGMM, C_GAUSSWISH, C_NORMGAMMA, DGMC, DGMM, F32, F64, GMC, SGMC, VDP, W_DIRICHLET,
                      W_GDIRICHLET, W_STICKBREAK, CudaError, DomainError, FreeEnergyError, InvalidArgument)

PRIORVAL = 1.0        # include/libcluster.h:122
SPLITITER = 15        # :124

_MODEL_CLUSTER = {VDP: C_GAUSSWISH, BGMM: C_GAUSSWISH, DGMM: C_NORMGAMMA, GMC: C_GAUSSWISH,
                  SGMC: C_GAUSSWISH, DGMC: C_NORMGAMMA}


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _as_matrix(X):
    """Accept C- or F-ordered float64 matrices without copying (Eigen row-/col-major)."""
    X = np.asarray(X)
    if X.ndim != 2:
        raise InvalidArgument("observations must be a 2-D array")
    if X.dtype != np.float64:
        X = X.astype(np.float64)
    if X.flags.c_contiguous:
        return X, nat.ROW_MAJOR, X.shape[1]
    if X.flags.f_contiguous:
        return X, nat.COL_MAJOR, X.shape[0]
    X = np.ascontiguousarray(X)
    return X, nat.ROW_MAJOR, X.shape[1]


class Engine:
    """One GPU's share of a fit: resident observations + responsibilities + posteriors."""

    def __init__(self, device=0, precision=F32):
        self._h = C.c_void_p()
        nat.check(nat.lib().lcb_create(C.byref(self._h), device, precision))
        self.precision = precision
        self._keep = None
        self._model = None
        self._D = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            nat.lib().lcb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- data ------------------------------------------------------------
    def set_data(self, X):
        groups = [X] if isinstance(X, np.ndarray) else list(X)
        mats = [_as_matrix(g) for g in groups]
        D = mats[0][0].shape[1]
        layout = mats[0][1]
        fixed = []
        for (m, lay, ld) in mats:
            if m.shape[1] != D:
                raise InvalidArgument("X dimensions are inconsistent between groups!")
            if lay != layout:
                m = np.ascontiguousarray(m) if layout == nat.ROW_MAJOR else np.asfortranarray(m)
                ld = m.shape[1] if layout == nat.ROW_MAJOR else m.shape[0]
            fixed.append((m, ld))
        J = len(fixed)
        ptrs = (C.POINTER(C.c_double) * J)(*[_dp(m) for m, _ in fixed])
        Nj = np.array([m.shape[0] for m, _ in fixed], dtype=np.int64)
        ld = np.array([max(l, 1) for _, l in fixed], dtype=np.int64)
        nat.check(nat.lib().lcb_set_data(self._h, J, ptrs, Nj.ctypes.data_as(C.POINTER(C.c_int64)), D,
                                         ld.ctypes.data_as(C.POINTER(C.c_int64)), layout))
        self._D = D
        self._Nj = Nj

    def set_data_device(self, X_dev_ptr, N, D, ld, gid_dev_ptr=None, J=1):
        """Adopt (copy) a row-major fp32 matrix already on this device, e.g. torch_tensor.data_ptr()."""
        nat.check(nat.lib().lcb_set_data_device_f32(self._h, C.c_void_p(X_dev_ptr), N, D, ld,
                                                    C.c_void_p(gid_dev_ptr) if gid_dev_ptr else None, J))
        self._D = D
        self._Nj = np.array([nat.lib().lcb_num_rows(self._h, j) for j in range(J)], dtype=np.int64)

    # ---- fits ------------------------------------------------------------
    def learn(self, model, prior=PRIORVAL, weight_prior=-1.0, maxclusters=-1, sparse=False, verbose=False,
              nthreads=1):
        F = C.c_double()
        K = C.c_int()
        self._model = model
        nat.check(nat.lib().lcb_learn(self._h, model, prior, weight_prior, maxclusters, int(sparse), int(verbose),
                                      nthreads, C.byref(F), C.byref(K)))
        return F.value

    def model_init(self, model, prior=PRIORVAL, weight_prior=-1.0, sparse=False):
        self._model = model
        nat.check(nat.lib().lcb_model_init(self._h, model, prior, weight_prior, int(sparse)))

    def set_qz(self, q0):
        q0 = np.ascontiguousarray(q0, dtype=np.float64)
        if q0.ndim != 2 or q0.shape[0] != int(self._Nj.sum()):
            raise InvalidArgument("qZ must be [N x K] over all rows")
        nat.check(nat.lib().lcb_set_qz(self._h, _dp(q0), q0.shape[1]))

    def set_labels_device(self, labels_dev_ptr, K):
        nat.check(nat.lib().lcb_set_labels_device(self._h, C.c_void_p(labels_dev_ptr), K))

    def vbem(self, maxit=-1):
        F = C.c_double()
        it = C.c_int()
        nat.check(nat.lib().lcb_vbem(self._h, maxit, C.byref(F), C.byref(it)))
        return F.value, it.value

    def vbem_step(self):
        F = C.c_double()
        nat.check(nat.lib().lcb_vbem_step(self._h, C.byref(F)))
        return F.value

    def step_timing(self):
        out = np.zeros(4)
        nat.check(nat.lib().lcb_get_step_timing(self._h, _dp(out)))
        return dict(sstat_ms=out[0], estep_ms=out[1], step_ms=out[2], launches=int(out[3]))

    def step_counts(self):
        """Launches, collectives and host synchronisations of the last vbem_step; device_mstep tells whether the
        posterior updates ran on the GPU (default) or on the host (LCB_HOST_MSTEP=1)."""
        out = np.zeros(4)
        nat.check(nat.lib().lcb_get_step_counts(self._h, _dp(out)))
        return dict(launches=int(out[0]), collectives=int(out[1]), host_syncs=int(out[2]), device_mstep=bool(out[3]))

    def estep_detail(self):
        """Break-down of the last tensor-core E pass (lcb_get_estep_detail)."""
        out = np.zeros(8)
        nat.check(nat.lib().lcb_get_estep_detail(self._h, _dp(out)))
        return dict(coarse_ms=out[0], lists_ms=out[1], refine_ms=out[2], finalize_ms=out[3], pairs=int(out[4]),
                    path=int(out[5]), items=int(out[6]))

    @property
    def stream(self):
        return nat.lib().lcb_stream(self._h)

    # ---- results -----------------------------------------------------------
    @property
    def K(self):
        return nat.lib().lcb_num_clusters(self._h)

    @property
    def J(self):
        return nat.lib().lcb_num_groups(self._h)

    def qZ(self, j=None, order="C", out=None):
        """qZ of group j as float64 [N_j x K].  `out`: an existing array of that shape and order to fill (a page-locked
        row-major one is written by DMA straight from the device)."""
        if j is None:
            return [self.qZ(g, order) for g in range(self.J)]
        Nj = int(nat.lib().lcb_num_rows(self._h, j))
        K = self._qcols()
        if out is None:
            out = np.empty((Nj, K), order=order)
        elif out.shape != (Nj, K) or out.dtype != np.float64 or not (out.flags.c_contiguous if order == "C" else out.flags.f_contiguous):
            raise InvalidArgument("qZ: out must be a float64 array of shape (%d, %d) in order %s" % (Nj, K, order))
        lay = nat.ROW_MAJOR if order == "C" else nat.COL_MAJOR
        nat.check(nat.lib().lcb_get_qz(self._h, j, _dp(out), max(K if order == "C" else Nj, 1), lay))
        return out

    def _qcols(self):
        return max(self.K, 1)

    def group_weights(self, j=0):
        K = max(self.K, 1)
        Nk = np.zeros(K)
        e = np.zeros(K)
        f = C.c_double()
        nat.check(nat.lib().lcb_get_group_weights(self._h, j, _dp(Nk), _dp(e), C.byref(f)))
        return Nk, e, f.value

    def cluster(self, k):
        D = self._D
        S = D * D if _MODEL_CLUSTER[self._model] == C_GAUSSWISH else D
        N_s, N, f = C.c_double(), C.c_double(), C.c_double()
        x_s, mean = np.zeros(D), np.zeros(D)
        xx_s, cov = np.zeros(S), np.zeros(S)
        nat.check(nat.lib().lcb_get_cluster(self._h, k, C.byref(N_s), _dp(x_s), _dp(xx_s), C.byref(N), _dp(mean),
                                            _dp(cov), C.byref(f)))
        shp = (D, D) if S == D * D else (D,)
        return dict(N_s=N_s.value, x_s=x_s, xx_s=xx_s.reshape(shp), N=N.value, mean=mean, cov=cov.reshape(shp),
                    fenergy=f.value)

    def trace(self):
        n = nat.lib().lcb_trace_len(self._h)
        F = np.zeros(n)
        K = np.zeros(n, dtype=np.int32)
        if n:
            nat.check(nat.lib().lcb_get_trace(self._h, _dp(F), K.ctypes.data_as(C.POINTER(C.c_int))))
        return F, K

    # ---- multi-GPU ---------------------------------------------------------
    def comm_init_nccl(self, unique_id, rank, world):
        nat.check(nat.lib().lcb_comm_init_nccl(self._h, unique_id, rank, world))

    def comm_init_host(self, fn, rank, world):
        """fn(numpy float64 view) must sum the buffer across ranks in place."""
        def _cb(buf, count, ctx):
            try:
                fn(np.ctypeslib.as_array(buf, shape=(count,)))
                return 0
            except Exception:
                return 1
        self._keep = nat.ALLREDUCE_FN(_cb)
        nat.check(nat.lib().lcb_comm_init_host(self._h, self._keep, None, rank, world))


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    nat.check(nat.lib().lcb_nccl_unique_id(buf))
    return buf.raw


# --------------------------------------------------------------------------
# learnXXX: python/libclusterpy.cpp:135-241 return conventions
# --------------------------------------------------------------------------
_DEFAULT = {"engine": None}


def default_engine(device=0, precision=F32):
    e = _DEFAULT["engine"]
    if e is None or e.precision != precision or e._h is None:
        e = Engine(device, precision)
        _DEFAULT["engine"] = e
    return e


def _fit(model, X, prior, weight_prior, maxclusters, sparse, verbose, nthreads, engine, precision):
    eng = engine if engine is not None else default_engine(precision=precision)
    eng.set_data(X)
    f = eng.learn(model, prior, weight_prior, maxclusters, sparse, verbose, nthreads)
    K = eng.K
    cl = [eng.cluster(k) for k in range(K)]
    mu = [c["mean"] for c in cl]
    cov = [c["cov"] for c in cl]
    w = [np.exp(eng.group_weights(j)[1]) for j in range(eng.J)]
    qZ = eng.qZ()
    return f, qZ, w, mu, cov


def learnVDP(X, prior=PRIORVAL, maxclusters=-1, verbose=False, nthreads=1, *, engine=None, precision=F32):
    """include/libcluster.h:177; returns (f, qZ, w, mu, cov) like libclusterpy.learnVDP."""
    f, qZ, w, mu, cov = _fit(VDP, X, prior, -1.0, maxclusters, False, verbose, nthreads, engine, precision)
    return f, qZ[0], w[0], mu, cov


def learnBGMM(X, prior=PRIORVAL, maxclusters=-1, verbose=False, nthreads=1, *, engine=None, precision=F32):
    """include/libcluster.h:218; returns (f, qZ, w, mu, cov)."""
    f, qZ, w, mu, cov = _fit(BGMM, X, prior, -1.0, maxclusters, False, verbose, nthreads, engine, precision)
    return f, qZ[0], w[0], mu, cov


def learnDGMM(X, prior=PRIORVAL, maxclusters=-1, verbose=False, nthreads=1, *, engine=None, precision=F32):
    """include/libcluster.h:262 (diagonal covariances: cov entries are D-vectors)."""
    f, qZ, w, mu, cov = _fit(DGMM, X, prior, -1.0, maxclusters, False, verbose, nthreads, engine, precision)
    return f, qZ[0], w[0], mu, cov


def learnGMC(X, prior=PRIORVAL, maxclusters=-1, sparse=False, verbose=False, nthreads=1, *, engine=None,
             precision=F32):
    """include/libcluster.h:356; X is a list of [N_j x D] arrays; qZ and w are lists over groups."""
    return _fit(GMC, X, prior, -1.0, maxclusters, sparse, verbose, nthreads, engine, precision)


def learnSGMC(X, prior=PRIORVAL, maxclusters=-1, sparse=False, verbose=False, nthreads=1, *, engine=None,
              precision=F32):
    """include/libcluster.h:409."""
    return _fit(SGMC, X, prior, -1.0, maxclusters, sparse, verbose, nthreads, engine, precision)


def learnDGMC(X, prior=PRIORVAL, maxclusters=-1, sparse=False, verbose=False, nthreads=1, *, engine=None,
              precision=F32):
    """include/libcluster.h:462."""
    return _fit(DGMC, X, prior, -1.0, maxclusters, sparse, verbose, nthreads, engine, precision)


# --------------------------------------------------------------------------
# operator classes: include/distributions.h
# --------------------------------------------------------------------------
class _WeightDist:
    _kind = None

    def __init__(self, prior=None):
        if prior is not None and prior <= 0:
            raise InvalidArgument("Concentration parameter has to be > 0!" if self._kind != W_DIRICHLET
                                  else "Alpha prior must be > 0!")
        self._h = C.c_void_p()
        nat.check(nat.lib().lcb_weights_create(C.byref(self._h), self._kind, -1.0 if prior is None else prior))

    def __del__(self):
        if getattr(self, "_h", None):
            nat.lib().lcb_weights_destroy(self._h)
            self._h = None

    def update(self, Nk):
        Nk = np.ascontiguousarray(Nk, dtype=np.float64)
        nat.check(nat.lib().lcb_weights_update(self._h, _dp(Nk), Nk.size))

    def Elogweight(self):
        out = np.zeros(nat.lib().lcb_weights_size(self._h))
        nat.check(nat.lib().lcb_weights_elogweight(self._h, _dp(out)))
        return out

    def getNk(self):
        out = np.zeros(nat.lib().lcb_weights_size(self._h))
        nat.check(nat.lib().lcb_weights_getnk(self._h, _dp(out)))
        return out

    def fenergy(self):
        return nat.lib().lcb_weights_fenergy(self._h)


class StickBreak(_WeightDist):
    _kind = W_STICKBREAK


class GDirichlet(_WeightDist):
    _kind = W_GDIRICHLET


class Dirichlet(_WeightDist):
    _kind = W_DIRICHLET


class _ClusterDist:
    _kind = None

    def __init__(self, clustwidth, D, engine=None, precision=F32):
        self._h = C.c_void_p()
        nat.check(nat.lib().lcb_cluster_create(C.byref(self._h), self._kind, clustwidth, D))
        self.D = D
        self._engine = engine
        self._precision = precision

    def _eng(self):
        return self._engine if self._engine is not None else default_engine(precision=self._precision)

    def __del__(self):
        if getattr(self, "_h", None):
            nat.lib().lcb_cluster_destroy(self._h)
            self._h = None

    def addobs(self, qZk, X):
        X, lay, ld = _as_matrix(X)
        qZk = np.ascontiguousarray(qZk, dtype=np.float64).ravel()
        if X.shape[1] != self.D:
            raise InvalidArgument("Mismatched dims. of cluster params and obs.!")
        if qZk.size != X.shape[0]:
            raise InvalidArgument("qZk and X ar not the same length!")
        nat.check(nat.lib().lcb_cluster_addobs(self._eng()._h, self._h, _dp(qZk), _dp(X), X.shape[0], max(ld, 1), lay))

    def update(self):
        nat.check(nat.lib().lcb_cluster_update(self._h))

    def clearobs(self):
        nat.check(nat.lib().lcb_cluster_clearobs(self._h))

    def Eloglike(self, X):
        X, lay, ld = _as_matrix(X)
        if X.shape[1] != self.D:
            raise InvalidArgument("Arguments do not have the same dimensionality")
        out = np.zeros(X.shape[0])
        nat.check(nat.lib().lcb_cluster_eloglike(self._eng()._h, self._h, _dp(X), X.shape[0], max(ld, 1), lay, _dp(out)))
        return out

    def splitobs(self, X):
        X, lay, ld = _as_matrix(X)
        out = np.zeros(X.shape[0], dtype=np.uint8)
        nat.check(nat.lib().lcb_cluster_splitobs(self._eng()._h, self._h, _dp(X), X.shape[0], max(ld, 1), lay,
                                                 out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out.astype(bool)

    def fenergy(self):
        return nat.lib().lcb_cluster_fenergy(self._h)

    def getN(self):
        return nat.lib().lcb_cluster_getn(self._h)

    def getprior(self):
        return nat.lib().lcb_cluster_getprior(self._h)

    def getmean(self):
        out = np.zeros(self.D)
        nat.check(nat.lib().lcb_cluster_getmean(self._h, _dp(out)))
        return out

    def _S(self):
        return self.D * self.D if self._kind == C_GAUSSWISH else self.D

    def getcov(self):
        out = np.zeros(self._S())
        nat.check(nat.lib().lcb_cluster_getcov(self._h, _dp(out)))
        return out.reshape((self.D, self.D)) if self._kind == C_GAUSSWISH else out

    def get_stats(self):
        N = C.c_double()
        xs, xxs = np.zeros(self.D), np.zeros(self._S())
        nat.check(nat.lib().lcb_cluster_get_stats(self._h, C.byref(N), _dp(xs), _dp(xxs)))
        return N.value, xs, (xxs.reshape((self.D, self.D)) if self._kind == C_GAUSSWISH else xxs)

    def set_stats(self, N_s, x_s, xx_s):
        x_s = np.ascontiguousarray(x_s, dtype=np.float64)
        xx_s = np.ascontiguousarray(xx_s, dtype=np.float64)
        nat.check(nat.lib().lcb_cluster_set_stats(self._h, float(N_s), _dp(x_s), _dp(xx_s)))


class GaussWish(_ClusterDist):
    _kind = C_GAUSSWISH


class NormGamma(_ClusterDist):
    _kind = C_NORMGAMMA


# ---- host-only iteration pieces (multi-rank tests) --------------------------
def packed_len(model, J, K, D):
    return int(nat.lib().lcb_packed_len(model, J, K, D))


def host_mstep(model, packed, J, K, D, prior=PRIORVAL, weight_prior=-1.0):
    packed = np.ascontiguousarray(packed, dtype=np.float64)
    S = D * D if _MODEL_CLUSTER[model] == C_GAUSSWISH else D
    F = C.c_double()
    e = np.zeros((J, K))
    means = np.zeros((K, D))
    covs = np.zeros((K, S))
    nat.check(nat.lib().lcb_host_mstep(model, prior, weight_prior, J, K, D, _dp(packed), C.byref(F), _dp(e), _dp(means),
                                       _dp(covs)))
    return F.value, e, means, covs


def shard_rows(N, rank, world):
    b, e = C.c_int64(), C.c_int64()
    nat.lib().lcb_shard_rows(N, rank, world, C.byref(b), C.byref(e))
    return b.value, e.value
