"""Independent numpy/scipy mirror of the libcluster VB loop (second opinion).

TEST INFRASTRUCTURE ONLY (see oracle/vb_oracle.c header).  Written separately
from the C restatement, against the same reference lines, with scipy's
digamma / gammaln / cho_factor so that an error in either restatement's special
functions or linear algebra shows up as a disagreement between the two.
Citations are path:line under /root/reference.
"""
import numpy as np
from scipy.linalg import cho_factor, cho_solve, solve_triangular
from scipy.special import digamma, gammaln

CONVERGE = float(np.float32(1e-5))            # include/libcluster.h:125
FENGYDEL = CONVERGE / 10                      # :126
ZEROCUTOFF = float(np.float32(0.1))           # :127
SPLITITER = 15                                # :124
EIGCONTHRESH = float(np.float32(1.0e-8))      # src/probutils.cpp:39
MAXITER = 100                                 # :40


class InvalidArgument(ValueError):
    pass


class FreeEnergyIncrease(RuntimeError):
    pass


# ---------------------------------------------------------------- weights ---
class StickBreak:                              # src/distributions.cpp:83-179
    def __init__(self, concentration=None):
        if concentration is not None and concentration <= 0:
            raise InvalidArgument("Concentration parameter has to be > 0!")
        self.a1p = 1.0 if concentration is None else float(concentration)
        self.a2p = 1.0
        self.Fp = gammaln(self.a1p) + gammaln(self.a2p) - gammaln(self.a1p + self.a2p)
        self.Nk = np.zeros(1)
        self.Elogpi = np.zeros(1)

    def _order(self, Nk):
        return np.argsort(-Nk, kind="stable")  # :146 (std::sort; ties impl.-defined)

    def update(self, Nk):
        Nk = np.asarray(Nk, dtype=float)
        K = Nk.size
        self.Nk = Nk.copy()
        self.a1 = self.a1p + Nk
        self.ord = self._order(Nk)
        self.a2 = np.zeros(K)
        self.Elogv = np.zeros(K)
        self.Elognv = np.zeros(K)
        self.Elogpi = np.zeros(K)
        N = Nk.sum()
        cum = 0.0
        cumE = 0.0
        for k in self.ord:
            cum += Nk[k]
            self.a2[k] = self.a2p + (N - cum)
            ps = digamma(self.a1[k] + self.a2[k])
            self.Elogv[k] = digamma(self.a1[k]) - ps
            self.Elognv[k] = digamma(self.a2[k]) - ps
            self.Elogpi[k] = self.Elogv[k] + cumE
            cumE += self.Elognv[k]

    def Elogweight(self):
        return self.Elogpi

    def getNk(self):
        return self.Nk

    def _terms(self, ks):
        a1, a2 = self.a1[ks], self.a2[ks]
        return (gammaln(a1 + a2) - gammaln(a1) - gammaln(a2)
                + (a1 - self.a1p) * self.Elogv[ks] + (a2 - self.a2p) * self.Elognv[ks]).sum()

    def fenergy(self):                         # :171-179
        K = self.a1.size
        return K * self.Fp + self._terms(np.arange(K))


class GDirichlet(StickBreak):                  # :186-215
    def update(self, Nk):
        super().update(Nk)
        sk = self.ord[-1]
        self.Elogpi[sk] -= self.Elogv[sk]
        self.Elogv[sk] = 0.0
        self.Elognv[sk] = 0.0

    def fenergy(self):
        K = self.ord.size
        return (K - 1) * self.Fp + (self._terms(self.ord[:-1]) if K > 1 else 0.0)


class Dirichlet:                               # :222-266
    def __init__(self, alpha=None):
        if alpha is not None and alpha <= 0:
            raise InvalidArgument("Alpha prior must be > 0!")
        self.ap = 1.0 if alpha is None else float(alpha)
        self.Nk = np.zeros(1)
        self.Elogpi = np.zeros(1)

    def update(self, Nk):
        Nk = np.asarray(Nk, dtype=float)
        self.Nk = Nk.copy()
        self.alpha = self.ap + Nk
        self.Elogpi = digamma(self.alpha) - digamma(self.alpha.sum())

    def Elogweight(self):
        return self.Elogpi

    def getNk(self):
        return self.Nk

    def fenergy(self):
        K = self.alpha.size
        return (gammaln(self.alpha.sum()) - (self.ap - 1) * self.Elogpi.sum()
                + ((self.alpha - 1) * self.Elogpi - gammaln(self.alpha)).sum()
                - gammaln(K * self.ap) + K * gammaln(self.ap))


# --------------------------------------------------------------- clusters ---
def eigpower(A):                               # src/probutils.cpp:153-186
    D = A.shape[0]
    if D == 1:
        return np.ones(1)
    v = np.linspace(-1, 1, D)
    e = v / np.linalg.norm(v)
    vdist = np.inf
    i = 0
    while vdist > EIGCONTHRESH and i < MAXITER:
        o = e
        v = A @ o
        e = v / np.linalg.norm(v)
        vdist = np.linalg.norm(e - o)
        i += 1
    return e


class GaussWish:                               # src/distributions.cpp:273-399
    def __init__(self, clustwidth, D):
        if clustwidth <= 0:
            raise InvalidArgument("clustwidth must be > 0!")
        self.D, self.prior, self.N = D, float(clustwidth), 0.0
        self.nu_p, self.beta_p = float(D), 1.0
        self.m_p = np.zeros(D)
        self.iW_p = self.nu_p * self.prior * np.eye(D)
        self.logdW_p = -np.linalg.slogdet(self.iW_p)[1]
        l = np.arange(1, D + 1)
        self.F_p = gammaln((self.nu_p + 1 - l) / 2).sum()
        self.clearobs()

    def clearobs(self):
        D = self.D
        self.nu, self.beta, self.m = self.nu_p, self.beta_p, self.m_p.copy()
        self.iW, self.logdW = self.iW_p.copy(), self.logdW_p
        self.N_s, self.x_s, self.xx_s = 0.0, np.zeros(D), np.zeros((D, D))

    def addobs(self, qk, X):
        qX = qk[:, None] * X
        self.N_s += qk.sum()
        self.x_s = self.x_s + qX.sum(0)
        self.xx_s = self.xx_s + qX.T @ X

    def update(self):
        xk = self.x_s / self.N_s if self.N_s > 0 else np.zeros(self.D)
        Sk = self.xx_s - np.outer(xk, self.x_s)
        d = xk - self.m_p
        self.N = self.N_s
        self.nu = self.nu_p + self.N
        self.beta = self.beta_p + self.N
        self.m = (self.beta_p * self.m_p + self.x_s) / self.beta
        self.iW = self.iW_p + Sk + (self.beta_p * self.N / self.beta) * np.outer(d, d)
        sign, ld = np.linalg.slogdet(self.iW)
        if sign <= 0:
            raise np.linalg.LinAlgError("Matrix A is not positive definite.")
        self.logdW = -ld

    def _maha(self, X):
        L = np.linalg.cholesky(self.iW)
        Z = solve_triangular(L, (X - self.m).T, lower=True)
        return (Z * Z).sum(0)

    def Eloglike(self, X):
        l = np.arange(1, self.D + 1)
        sumpsi = digamma((self.nu + 1 - l) / 2).sum()
        return 0.5 * (sumpsi + self.logdW - self.D * (1 / self.beta + np.log(np.pi))
                      - self.nu * self._maha(X))

    def splitobs(self, X):
        return ((X - self.m) @ eigpower(self.iW)) >= 0

    def fenergy(self):
        D = self.D
        l = np.arange(1, D + 1)
        sumpsi = digamma((self.nu + 1 - l) / 2).sum()
        cf = cho_factor(self.iW)
        tr = np.trace(cho_solve(cf, self.iW_p))
        dm = self.m - self.m_p
        mh = dm @ cho_solve(cf, dm)
        return (self.F_p + (D * (self.beta_p / self.beta - 1 - self.nu - np.log(self.beta_p / self.beta))
                            + self.nu * (tr + self.beta_p * mh)
                            + self.nu_p * (self.logdW_p - self.logdW) + self.N * sumpsi) / 2
                - gammaln((self.nu + 1 - l) / 2).sum())

    def getN(self):
        return self.N

    def getmean(self):
        return self.m

    def getcov(self):
        return self.iW / self.nu


class NormGamma:                               # src/distributions.cpp:406-517
    def __init__(self, clustwidth, D):
        if clustwidth <= 0:
            raise InvalidArgument("clustwidth must be > 0!")
        self.D, self.prior, self.N = D, float(clustwidth), 0.0
        self.nu_p, self.beta_p = 1.0, 1.0
        self.m_p = np.zeros(D)
        self.L_p = self.nu_p * self.prior * np.ones(D)
        self.logL_p = np.log(self.L_p).sum()
        self.clearobs()

    def clearobs(self):
        D = self.D
        self.nu, self.beta, self.m = self.nu_p, self.beta_p, self.m_p.copy()
        self.L, self.logL = self.L_p.copy(), self.logL_p
        self.N_s, self.x_s, self.xx_s = 0.0, np.zeros(D), np.zeros(D)

    def addobs(self, qk, X):
        qX = qk[:, None] * X
        self.N_s += qk.sum()
        self.x_s = self.x_s + qX.sum(0)
        self.xx_s = self.xx_s + (qX * X).sum(0)

    def update(self):
        xk, Sk = np.zeros(self.D), np.zeros(self.D)
        if self.N_s > 0:
            xk = self.x_s / self.N_s
            Sk = self.xx_s - self.x_s ** 2 / self.N_s
        self.N = self.N_s
        self.beta = self.beta_p + self.N
        self.nu = self.nu_p + self.N / 2
        self.m = (self.beta_p * self.m_p + self.x_s) / self.beta
        self.L = self.L_p + Sk / 2 + (self.beta_p * self.N / (2 * self.beta)) * (xk - self.m_p) ** 2
        if (self.L <= 0).any():
            raise InvalidArgument("Calc log(L): Variance is zero or less!")
        self.logL = np.log(self.L).sum()

    def Eloglike(self, X):
        dist = ((X - self.m) ** 2) @ (1.0 / self.L)
        return 0.5 * (self.D * (digamma(self.nu) - np.log(2 * np.pi) - 1 / self.beta)
                      - self.logL - self.nu * dist)

    def splitobs(self, X):
        e = int(np.argmax(self.L))
        return (X[:, e] - self.m[e]) >= 0

    def fenergy(self):
        D = self.D
        iL = 1.0 / self.L
        return (D * (gammaln(self.nu_p) - gammaln(self.nu) + self.N * digamma(self.nu) / 2 - self.nu)
                + (D // 2) * (np.log(self.beta) - np.log(self.beta_p) - 1 + self.beta_p / self.beta)
                + self.beta_p * self.nu / 2 * ((self.m - self.m_p) ** 2 @ iL)
                + self.nu_p * (self.logL - self.logL_p) + self.nu * (self.L_p @ iL))

    def getN(self):
        return self.N

    def getmean(self):
        return self.m

    def getcov(self):
        return self.L * self.nu                # include/distributions.h:375 (sic)


# ------------------------------------------------------------- algorithms ---
def _kful(K, sparse, Nk):
    if not sparse:
        return np.arange(K) if K > 1 else np.zeros(1, dtype=int)
    return np.nonzero(Nk >= ZEROCUTOFF)[0]


def vbem(X, qZ, weights, clusters, W, C, prior, maxit=-1, sparse=False, trace=None):
    """src/cluster.cpp:177-239.  X, qZ: lists over groups; mutates qZ, weights, clusters."""
    J, K, D = len(X), qZ[0].shape[1], X[0].shape[1]
    while len(weights) < J:
        weights.append(W())
    del weights[J:]
    while len(clusters) < K:
        clusters.append(C(prior, D))
    del clusters[K:]
    F = np.finfo(float).max
    i = 0
    while True:
        Fold = F
        for c in clusters:
            c.clearobs()
        for j in range(J):                     # updateSS, :53-82
            Njk = qZ[j].sum(0)
            for k in _kful(K, sparse, Njk):
                clusters[k].addobs(qZ[j][:, k], X[j])
            weights[j].update(Njk)
        for c in clusters:
            c.update()
        Fz = 0.0
        for j in range(J):                     # vbexpectation, :91-138
            ElogZ = weights[j].Elogweight()
            ful = _kful(K, sparse, weights[j].getNk())
            lq = np.stack([ElogZ[k] + clusters[k].Eloglike(X[j]) for k in ful], axis=1) \
                if X[j].shape[0] else np.zeros((0, len(ful)))
            mx = lq.max(1) if lq.size else np.zeros(0)
            lz = np.log(np.exp(lq - mx[:, None]).sum(1)) + mx
            q = np.zeros((X[j].shape[0], K))
            q[:, ful] = np.exp(lq - lz[:, None])
            qZ[j] = q
            Fz += -lz.sum()
        F = sum(w.fenergy() for w in weights) + sum(c.fenergy() for c in clusters) + Fz
        if trace is not None:
            trace.append((F, K))
        if (F - Fold) / abs(Fold) > FENGYDEL:
            raise FreeEnergyIncrease("Free energy increase!")
        cont = abs((Fold - F) / Fold) > CONVERGE
        if cont:
            cont = (i < maxit) or (maxit < 0)
            i += 1
        if not cont:
            break
    return F


def prune_clusters(qZ, weights, clusters):     # src/cluster.cpp:505-552
    Nk = np.array([c.getN() for c in clusters])
    empty = Nk < ZEROCUTOFF
    if not empty.any():
        return False
    keep = np.nonzero(~empty)[0]
    for k in sorted(np.nonzero(empty)[0], reverse=True):
        del clusters[k]
    for j in range(len(qZ)):
        qZ[j] = qZ[j][:, keep].copy()
        weights[j].update(qZ[j].sum(0))
    return True


def split_gr(X, weights, clusters, qZ, tally, F, maxclusters, sparse, W, C, trace=None):
    """src/cluster.cpp:367-495; returns True when a split was accepted (qZ replaced)."""
    J, K = len(X), len(clusters)
    if K >= maxclusters >= 0:
        return False
    while len(tally) < K:
        tally.append(0)
    del tally[K:]
    Fk = np.array([c.fenergy() for c in clusters])
    for j in range(J):
        lp = weights[j].Elogweight()
        for k in range(K):
            Fk[k] -= qZ[j][:, k] @ (lp[k] + clusters[k].Eloglike(X[j]))
    order = sorted(range(K), key=lambda k: (tally[k], -Fk[k]))   # comutils.h:60-68
    for k in order:
        tally[k] += 1
        if clusters[k].getN() < 4:
            continue
        mapidx = [np.nonzero(qZ[j][:, k] > 0.5)[0] for j in range(J)]
        Xk = [X[j][mapidx[j]] for j in range(J)]
        split = [clusters[k].splitobs(Xk[j]) if Xk[j].shape[0] else np.zeros(0, bool) for j in range(J)]
        qref = [np.stack([s.astype(float), (~s).astype(float)], 1) for s in split]
        scount = sum(int(s.sum()) for s in split)
        Mtot = sum(x.shape[0] for x in Xk)
        if scount < 2 or scount > Mtot - 2:
            continue
        wspl, cspl = [], []
        vbem(Xk, qref, wspl, cspl, W, C, clusters[0].prior, SPLITITER, sparse, trace)
        if any(c.getN() <= 1 for c in cspl):
            continue
        qaug = []
        for j in range(J):                     # auglabels, src/comutils.cpp:75-104
            qa = np.hstack([qZ[j], np.zeros((qZ[j].shape[0], 1))])
            rows = mapidx[j][qref[j][:, 1] > 0.5]
            qa[rows, K] = qZ[j][rows, k]
            qa[rows, k] = 0.0
            qaug.append(qa)
        Fsplit = vbem(X, qaug, wspl, cspl, W, C, clusters[0].prior, 1, sparse, trace)
        if any(c.getN() <= 1 for c in cspl):
            continue
        if Fsplit < F and abs((F - Fsplit) / F) > CONVERGE:
            qZ[:] = qaug
            tally[k] = 0
            return True
    return False


def cluster(X, W, C, prior=1.0, maxclusters=-1, sparse=False, weights=None):
    """src/cluster.cpp:564-629.  Returns (F, qZ, weights, clusters, trace)."""
    X = [np.asarray(x, dtype=float) for x in X]
    qZ = [np.ones((x.shape[0], 1)) for x in X]
    weights = [] if weights is None else weights
    clusters, tally, trace = [], [], []
    while True:
        F = vbem(X, qZ, weights, clusters, W, C, prior, -1, sparse, trace)
        prune_clusters(qZ, weights, clusters)
        if not split_gr(X, weights, clusters, qZ, tally, F, maxclusters, sparse, W, C, trace):
            break
    return F, qZ, weights, clusters, trace


MODELS = {
    "VDP": (StickBreak, GaussWish), "BGMM": (Dirichlet, GaussWish), "DGMM": (Dirichlet, NormGamma),
    "GMC": (GDirichlet, GaussWish), "SGMC": (Dirichlet, GaussWish), "DGMC": (GDirichlet, NormGamma),
}


def learn(model, X, prior=1.0, maxclusters=-1, sparse=False):
    W, C = MODELS[model]
    if isinstance(X, np.ndarray):
        X = [X]
    return cluster(X, W, C, prior, maxclusters, sparse)
