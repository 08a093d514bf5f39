"""GPU (-m gpu): the two-level E step of the tensor-core tier (D = 128, fp32 engine).

Level 1 computes every (row, cluster) distance with one fp16 product and a rigorous error bound and marks the
pairs that can matter; level 2 recomputes those at fp32-equivalent accuracy; level 3 is the row soft-max.  The
tests check each level against the dense tensor-core kernel (LCB_TC_TWO_LEVEL=0), the fp64 engine and the oracle.
"""
import os

import numpy as np
import pytest

import libcluster_b200 as lc
from conftest import make_blobs, soft_labels
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

D = 128
MARGIN = 24.0


def engine(two_level, stage=None, prec=lc.F32):
    os.environ["LCB_TC_TWO_LEVEL"] = "1" if two_level else "0"
    if stage:
        os.environ["LCB_TC_STAGE"] = stage
    try:
        return lc.Engine(0, prec)
    finally:
        os.environ.pop("LCB_TC_TWO_LEVEL", None)
        os.environ.pop("LCB_TC_STAGE", None)


def one_iteration(X, q0, model=lc.BGMM, sparse=False, **kw):
    eng = engine(**kw)
    eng.set_data(X)
    eng.model_init(model, sparse=sparse)
    eng.set_qz(q0)
    F, _ = eng.vbem(maxit=0)
    q = eng.qZ() if isinstance(X, list) else eng.qZ(0)
    if isinstance(q, list):
        q = np.concatenate(q, 0)
    det = eng.estep_detail()
    eng.close()
    return F, q, det


def softmax_rows(L):
    m = L.max(1, keepdims=True)
    e = np.exp(L - m)
    return e / e.sum(1, keepdims=True)


CASES = [  # N, K, spread (small spread = overlapping clusters = many candidates)
    (5000, 16, 4.0),
    (4099, 10, 1.0),
    (20000, 33, 3.0),
    (3000, 64, 6.0),
    (1500, 8, 0.3),
]


@pytest.mark.parametrize("N,K,spread", CASES)
def test_levels_against_dense_kernel_and_fp64(N, K, spread):
    X, z = make_blobs(N, D, K, seed=N + K, spread=spread)
    q0 = soft_labels(z, K, seed=K)
    F_d, q_d, det_d = one_iteration(X, q0, two_level=False)
    assert det_d["path"] == 0
    _, q64, _ = one_iteration(X, q0, two_level=False, prec=lc.F64)

    # level 1: UB for candidates, -inf for the rest
    _, ub, det1 = one_iteration(X, q0, two_level=True, stage="coarse")
    cand = np.isfinite(ub)
    assert det1["path"] == 1
    assert cand.any(1).all(), "every row keeps at least its best cluster"
    # a pair that is not a candidate is at least e^-margin below the row's best (fp64 responsibilities)
    with np.errstate(divide="ignore"):
        rel64 = np.log(q64) - np.log(q64.max(1, keepdims=True))
    assert (rel64[~cand] <= -MARGIN + 1e-2).all(), rel64[~cand].max()

    # level 2: exact logits for the candidates; their soft-max is the dense kernel's q
    _, lg, det2 = one_iteration(X, q0, two_level=True, stage="refine")
    assert (np.isfinite(lg) == cand).all()
    assert det2["pairs"] == cand.sum()
    L = np.where(cand, lg, -np.inf)
    assert np.abs(softmax_rows(L) - q_d).max() <= 2e-6
    # the level-1 upper bound really bounds the exact logit (up to the fp32 noise of the exact one)
    slack = ub[cand] - lg[cand]
    assert slack.min() >= -1e-2 * (1 + np.abs(lg[cand]).max() * 1e-5), slack.min()

    # all three levels
    F_t, q_t, det = one_iteration(X, q0, two_level=True)
    assert det["path"] in (1, 2)
    assert np.abs(q_t - q_d).max() <= 2e-6
    assert abs(F_t - F_d) <= 1e-7 * abs(F_d)
    assert np.allclose(q_t.sum(1), 1.0, atol=1e-5)
    assert np.abs(q_t - q64).max() <= 1e-5


def test_two_level_matches_oracle_over_iterations():
    N, K = 6000, 12
    X, z = make_blobs(N, D, K, seed=5, spread=3.0)
    q0 = soft_labels(z, K, seed=2)
    m = po.Model(po.BGMM, [X])
    m.vbem(q0, maxit=3)
    Fo, _ = m.trace()
    eng = engine(True)
    eng.set_data(X)
    eng.model_init(lc.BGMM)
    eng.set_qz(q0)
    eng.vbem(maxit=3)
    assert eng.estep_detail()["path"] == 1
    assert np.allclose(eng.trace()[0], Fo, rtol=1e-5)
    assert np.abs(eng.qZ(0) - m.qZ()).max() <= 1e-5
    eng.close()


@pytest.mark.parametrize("spread,lo,hi,tol3", [(3.5, 2.0, 12.0, 1e-5), (3.2, 4.0, 24.0, 2e-5)])
def test_headline_shape_against_oracle(spread, lo, hi, tol3):
    """The headline shape (D = 128, K = 64) directly against the oracle in the soft regime the benchmark mixture
    never visits: a broad cluster prior (clustwidth 10) on 4096 rows leaves several clusters in reach of every
    row (about 5 and 9 pairs per row with q > e^-24 in the oracle after the first iteration), so levels 2-3 and the S pass work on many
    pairs per row.  F of every iteration holds the stated 1e-5 and so does qZ after one iteration; after three
    iterations qZ is at 3e-6 in the first case and at about 1e-5 in the second (run-to-run: the statistics are summed
    with atomics), which is held to 2e-5."""
    N, K = 4096, 64
    X, z = make_blobs(N, D, K, seed=5, spread=spread)
    q0 = soft_labels(z, K, seed=1, noise=0.2)
    m = po.Model(po.BGMM, [X])
    m.vbem(q0, prior=10.0, maxit=0)
    q1 = m.qZ()                      # after the first iteration the assignments are still soft
    m = po.Model(po.BGMM, [X])
    m.vbem(q0, prior=10.0, maxit=2)
    Fo, _ = m.trace()
    qo = m.qZ()
    for maxit, qref in ((0, q1), (2, qo)):
        eng = engine(True)
        eng.set_data(X)
        eng.model_init(lc.BGMM, prior=10.0)
        eng.set_qz(q0)
        eng.vbem(maxit=maxit)
        det = eng.estep_detail()
        assert det["path"] == 1, det
        per_row = det["pairs"] / N
        assert per_row >= (qref > np.exp(-MARGIN)).sum(1).mean() - 1e-9   # the candidates are a superset
        if maxit == 0:
            assert lo <= per_row <= hi, per_row
        F = eng.trace()[0]
        assert len(F) == maxit + 1
        assert np.allclose(F, Fo[:maxit + 1], rtol=1e-5, atol=0)
        assert np.abs(eng.qZ(0) - qref).max() <= (1e-5 if maxit == 0 else tol3)
        if maxit == 2:
            assert np.allclose(eng.group_weights(0)[0], m.weights(0)[1], rtol=1e-5, atol=1e-3)
        eng.close()


def test_two_level_grouped_and_sparse():
    K = 9
    X, z = make_blobs(5000, D, K, seed=11, spread=3.0)
    order = np.argsort(z, kind="stable")
    X, z = X[order], z[order]
    groups = [X[:1800], X[1800:2100], X[2100:]]
    q0 = soft_labels(z, K, seed=4)
    for model, sparse in ((lc.GMC, False), (lc.SGMC, True)):
        F_d, q_d, _ = one_iteration(groups, q0, model=model, sparse=sparse, two_level=False)
        F_t, q_t, det = one_iteration(groups, q0, model=model, sparse=sparse, two_level=True)
        assert det["path"] == 1
        assert np.abs(q_t - q_d).max() <= 2e-6, (model, sparse)
        assert abs(F_t - F_d) <= 1e-7 * abs(F_d)


def test_two_level_gives_up_when_everything_is_a_candidate():
    # one blob labelled at random: every cluster posterior covers every point
    rng = np.random.default_rng(0)
    N, K = 4000, 8
    X = rng.normal(size=(N, D))
    q0 = soft_labels(rng.integers(0, K, N), K, seed=1, noise=0.5)
    F_d, q_d, _ = one_iteration(X, q0, two_level=False)
    F_t, q_t, det = one_iteration(X, q0, two_level=True)
    assert det["path"] == 2 and det["pairs"] > 0.4 * K * N
    # the same dense kernel ran in both engines; what differs is the summation order of the fp32 statistics pass
    # (atomic list positions), which this maximally overlapping case amplifies
    assert np.abs(q_t - q_d).max() <= 1e-4 and abs(F_t - F_d) <= 1e-7 * abs(F_d)


def test_two_level_full_fit_with_splits():
    X, _ = make_blobs(6000, D, 10, seed=21, spread=4.0)
    res = {}
    for two in (False, True):
        eng = engine(two)
        eng.set_data(X)
        eng.learn(lc.BGMM)
        res[two] = (eng.K, eng.trace()[0], eng.qZ(0))
        eng.close()
    assert res[True][0] == res[False][0]
    assert len(res[True][1]) == len(res[False][1])
    assert np.allclose(res[True][1], res[False][1], rtol=1e-6)
    assert np.abs(res[True][2] - res[False][2]).max() <= 1e-5


def test_two_level_size_independent_properties_at_scale():
    """N = 600k x 128, K = 64 resident on the device (the bench's generator at reduced N; the CPU oracle cannot reach
    this size): over three VB iterations the two-level pass and the dense kernel give the same F (relative 1e-8),
    every row keeps at least one candidate, rows of qZ sum to one and q is non-negative, and F does not increase."""
    import torch

    N, K = 600_000, 64
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(7)
    mu = torch.rand(K, D, device=dev, generator=g) * 20 - 10
    z = torch.randint(0, K, (N,), device=dev, generator=g).to(torch.int32)
    A = torch.randn(K, D, D, device=dev, generator=g)
    Lc = torch.linalg.cholesky(A @ A.transpose(1, 2) / D + 0.5 * torch.eye(D, device=dev))
    X = torch.empty(N, D, device=dev)
    for k in range(K):
        idx = (z == k).nonzero().squeeze(1)
        X[idx] = mu[k] + torch.randn(idx.numel(), D, device=dev, generator=g) @ Lc[k].T
    torch.cuda.synchronize()
    out = {}
    for two in (False, True):
        eng = engine(two)
        eng.set_data_device(X.data_ptr(), N, D, D)
        eng.model_init(lc.BGMM)
        eng.set_labels_device(z.data_ptr(), K)
        Fs, det = [], None
        for _ in range(3):
            Fs.append(eng.vbem_step())
            det = eng.estep_detail()
        q = eng.qZ(0)
        eng.close()
        out[two] = (Fs, q, det)
    Fd, qd, _ = out[False]
    Ft, qt, det = out[True]
    assert det["path"] == 1 and det["pairs"] >= N
    for a, b in zip(Fd, Ft):
        assert abs(a - b) <= 1e-8 * abs(a)
    assert all(Ft[i + 1] <= Ft[i] + 1e-8 * abs(Ft[i]) for i in range(len(Ft) - 1))   # cluster.cpp:229-230, fp32 noise
    assert qt.min() >= 0.0 and np.abs(qt.sum(1) - 1.0).max() <= 1e-5
    assert np.abs(qt - qd).max() <= 2e-6
