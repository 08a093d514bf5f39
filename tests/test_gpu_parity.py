"""GPU (-m gpu): the CUDA engine against the oracle, through the C ABI.

Tolerances (north_star: fp64 -> fp32, 1e-5 on qZ and F):
  LCB_F64 engine : F rel 1e-9,  qZ abs 1e-8   (same arithmetic, other summation order)
  LCB_F32 engine : F rel 1e-5,  qZ abs 1e-5   (the measured path)
"""
import numpy as np
import pytest

import libcluster_b200 as lc
from conftest import golden, make_blobs, soft_labels
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

TOL = {lc.F64: dict(F=1e-9, q=1e-8), lc.F32: dict(F=1e-5, q=1e-5)}
PRECS = [lc.F64, lc.F32]


@pytest.fixture(scope="module")
def engines():
    e = {p: lc.Engine(0, p) for p in PRECS}
    yield e
    for x in e.values():
        x.close()


def _compare_fit(eng, m, prec, check_trace=True):
    t = TOL[prec]
    assert eng.K == m.K
    Fe, Ke = eng.trace()
    Fo, Ko = m.trace()
    if check_trace:
        assert len(Fe) == len(Fo), (len(Fe), len(Fo))
        assert np.array_equal(Ke, Ko)
        assert np.allclose(Fe, Fo, rtol=t["F"], atol=0)
    q = np.concatenate(eng.qZ(), 0)
    qo = m.qZ()
    assert q.shape == qo.shape
    assert np.abs(q - qo).max() <= t["q"]
    assert np.allclose(q.sum(1), 1.0, atol=1e-5)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("model,omodel,N,D,K", [
    (lc.BGMM, po.BGMM, 5000, 2, 3),
    (lc.BGMM, po.BGMM, 3000, 5, 1),
    (lc.VDP, po.VDP, 4000, 16, 6),
    (lc.BGMM, po.BGMM, 4096, 64, 8),
    (lc.VDP, po.VDP, 3000, 128, 5),
    (lc.BGMM, po.BGMM, 2000, 19, 4),     # D not a multiple of anything
    (lc.DGMM, po.DGMM, 5000, 7, 4),
    (lc.DGMM, po.DGMM, 3000, 96, 6),
])
def test_vbem_iterations_match_oracle(engines, prec, model, omodel, N, D, K):
    """vbem() (cluster.cpp:177-239) from the same soft labels: F of every iteration, final qZ, statistics."""
    diag = omodel == po.DGMM
    X, z = make_blobs(N, D, K, seed=N + D, spread=4.0, diag=diag)
    q0 = soft_labels(z, K, seed=D)
    m = po.Model(omodel, [X])
    m.vbem(q0, maxit=3)
    eng = engines[prec]
    eng.set_data(X)
    eng.model_init(model)
    eng.set_qz(q0)
    F, it = eng.vbem(maxit=3)
    assert it == len(m.trace()[0])
    _compare_fit(eng, m, prec)
    for k in range(K):
        ce, co = eng.cluster(k), m.cluster(k)
        assert ce["N"] == pytest.approx(co["N"], rel=10 * TOL[prec]["F"])
        assert np.allclose(ce["mean"], co["m"], rtol=0, atol=20 * TOL[prec]["q"] * (1 + np.abs(co["m"]).max()))


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("name,model,grouped", [
    ("bgmm_xcat", lc.BGMM, False), ("vdp_xcat", lc.VDP, False), ("dgmm_xcat", lc.DGMM, False),
    ("gmc_groups", lc.GMC, True), ("sgmc_groups", lc.SGMC, True), ("dgmc_groups", lc.DGMC, True)])
def test_learn_on_reference_testdata_matches_golden(engines, testdata, prec, name, model, grouped):
    """BASELINE config 1: learnXXX on test/testdata.h makeXdata -- K, F, qZ, whole F trace."""
    X, _ = testdata
    groups = list(X) if grouped else np.concatenate(list(X), 0)
    g = golden(name)
    eng = engines[prec]
    eng.set_data(groups)
    F = eng.learn(model)
    t = TOL[prec]
    assert eng.K == int(g["K"])
    assert F == pytest.approx(float(g["F"]), rel=t["F"])
    Fe, Ke = eng.trace()
    assert len(Fe) == len(g["trace_F"])
    assert np.array_equal(Ke, g["trace_K"])
    assert np.allclose(Fe, g["trace_F"], rtol=t["F"])
    q = np.concatenate(eng.qZ(), 0)
    assert np.abs(q - g["qZ"]).max() <= t["q"]
    means = np.stack([eng.cluster(k)["mean"] for k in range(eng.K)])
    assert np.allclose(means, g["means"], atol=1e-4)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("model,omodel,D,Kt,diag", [(lc.BGMM, po.BGMM, 3, 4, False), (lc.VDP, po.VDP, 8, 5, False),
                                                   (lc.DGMM, po.DGMM, 6, 4, True)])
def test_learn_with_splits_matches_oracle(engines, prec, model, omodel, D, Kt, diag):
    """cluster() from K=1 through greedy splits (cluster.cpp:564-629, :367-495) on well separated blobs."""
    X, _ = make_blobs(2500, D, Kt, seed=40 + D, spread=8.0, diag=diag)
    m = po.Model(omodel, [X])
    m.learn()
    eng = engines[prec]
    eng.set_data(X)
    eng.learn(model)
    _compare_fit(eng, m, prec)


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("model,omodel,sparse", [(lc.GMC, po.GMC, False), (lc.SGMC, po.SGMC, False),
                                                 (lc.DGMC, po.DGMC, False), (lc.GMC, po.GMC, True)])
def test_grouped_models_match_oracle(engines, prec, model, omodel, sparse):
    """learnGMC/SGMC/DGMC (cluster.cpp:763-831): per-group weights, shared clusters, optional sparse updates."""
    diag = omodel == po.DGMC
    X, z = make_blobs(3000, 4, 5, seed=77, spread=9.0, diag=diag)
    order = np.argsort(z, kind="stable")
    X = X[order]
    # 6 groups of unequal size, each seeing a different subset of the clusters; one tiny group
    cuts = [0, 700, 1300, 1310, 2000, 2600, 3000]
    groups = [X[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    m = po.Model(omodel, groups)
    m.learn(sparse=sparse)
    eng = engines[prec]
    eng.set_data(groups)
    eng.learn(model, sparse=sparse)
    _compare_fit(eng, m, prec, check_trace=not (sparse and prec == lc.F32))
    for j in range(len(groups)):
        Nk, elw, fw = eng.group_weights(j)
        eo, no = m.weights(j)
        assert np.allclose(Nk, no, atol=1e-2 if prec == lc.F32 else 1e-6)
        assert np.allclose(elw, eo, atol=1e-4 if prec == lc.F32 else 1e-8)


@pytest.mark.parametrize("prec", PRECS)
def test_maxclusters_and_column_major_input(engines, prec):
    X, _ = make_blobs(1500, 3, 4, seed=9, spread=8.0)
    m = po.Model(po.BGMM, [X])
    m.learn(maxclusters=2)
    eng = engines[prec]
    eng.set_data(np.asfortranarray(X))          # Eigen's default storage order
    eng.learn(lc.BGMM, maxclusters=2)
    assert eng.K == m.K <= 2
    _compare_fit(eng, m, prec)
    qf = eng.qZ(0, order="F")
    assert np.abs(qf - m.qZ()).max() <= TOL[prec]["q"]


def test_python_learn_wrappers_return_reference_tuples(testdata):
    """python/libclusterpy.cpp:156-157,211-212 -- (f, qZ, w, mu, cov)."""
    X, _ = testdata
    Xcat = np.concatenate(list(X), 0)
    f, qZ, w, mu, cov = lc.learnBGMM(Xcat)
    g = golden("bgmm_xcat")
    assert f == pytest.approx(float(g["F"]), rel=1e-5)
    assert qZ.shape == (120, 3) and len(mu) == 3 and cov[0].shape == (2, 2)
    assert np.allclose(w, np.exp(g["Elogweight"][0]), atol=1e-4)
    f, qZ, w, mu, cov = lc.learnGMC(list(X))
    assert len(qZ) == 12 and len(w) == 12 and qZ[0].shape[0] == 10
    f, qZ, w, mu, cov = lc.learnDGMM(Xcat)
    assert cov[0].shape == (2,)
    f2, *_ = lc.learnVDP(Xcat, 1.0, -1, False, 4)
    assert f2 == pytest.approx(float(golden("vdp_xcat")["F"]), rel=1e-5)


def test_error_behaviour_matches_reference():
    X, _ = make_blobs(50, 2, 1, seed=1)
    with pytest.raises(lc.InvalidArgument, match="at least one thread"):
        lc.learnBGMM(X, nthreads=0)                  # cluster.cpp:576-577
    with pytest.raises(lc.InvalidArgument, match="clustwidth"):
        lc.learnBGMM(X, prior=-1.0)                  # distributions.cpp:282-283
    with pytest.raises(lc.InvalidArgument):
        lc.learnGMC([X, X[:, :1]])                   # inconsistent D between groups
    eng = lc.Engine(0, lc.F32)
    with pytest.raises(lc.InvalidArgument):
        eng.learn(lc.BGMM)                           # no data
    eng.close()


@pytest.mark.parametrize("prec", PRECS)
def test_edge_cases(engines, prec):
    eng = engines[prec]
    # fewer points than a tile, one point per cluster impossible to split (getN < 4)
    X = np.array([[0.0, 0.0], [0.1, 0.0], [0.0, 0.1]])
    m = po.Model(po.BGMM, [X]); m.learn()
    eng.set_data(X); eng.learn(lc.BGMM)
    _compare_fit(eng, m, prec)
    # a group with no rows at all (ragged input)
    Xb, _ = make_blobs(600, 2, 2, seed=4, spread=9.0)
    groups = [Xb[:300], Xb[:0], Xb[300:]]
    m = po.Model(po.GMC, groups); m.learn()
    eng.set_data(groups); eng.learn(lc.GMC)
    _compare_fit(eng, m, prec)
    assert eng.qZ(1).shape[0] == 0
    # D = 1
    X1 = np.concatenate([np.random.default_rng(0).normal(-5, 1, 400), np.random.default_rng(1).normal(5, 1, 400)])[:, None]
    m = po.Model(po.VDP, [X1]); m.learn()
    eng.set_data(X1); eng.learn(lc.VDP)
    _compare_fit(eng, m, prec)


def test_full_size_properties_f32_vs_f64():
    """BASELINE-scale shape (D=64, K=32) at a size the oracle cannot reach: size-independent properties --
    rows of qZ sum to 1, sum_k N_k = N, F32 and F64 engines agree on F (1e-5) and qZ (1e-5)."""
    import torch
    N, D, K = 1 << 20, 64, 32
    gen = torch.Generator(device="cuda").manual_seed(1234)
    mu = (torch.rand(K, D, device="cuda", generator=gen) * 2 - 1) * 3.0
    z = torch.randint(0, K, (N,), device="cuda", generator=gen, dtype=torch.int32)
    X = mu[z.long()] + torch.randn(N, D, device="cuda", generator=gen)
    res = {}
    for prec in (lc.F32, lc.F64):
        eng = lc.Engine(0, prec)
        eng.set_data_device(X.data_ptr(), N, D, D)
        eng.model_init(lc.BGMM)
        eng.set_labels_device(z.data_ptr(), K)
        Fs = [eng.vbem_step() for _ in range(3)]
        Nk = eng.group_weights(0)[0]
        q = eng.qZ(0)
        res[prec] = (Fs, Nk, q)
        assert np.allclose(q.sum(1), 1.0, atol=2e-6)
        assert Nk.sum() == pytest.approx(N, rel=1e-9)
        # F never rises by more than the reference's own tolerance (cluster.cpp:229: FENGYDEL = 1e-6 relative)
        assert all((b - a) / abs(a) <= 1e-6 for a, b in zip(Fs[:-1], Fs[1:])), Fs
        eng.close()
    assert np.allclose(res[lc.F32][0], res[lc.F64][0], rtol=1e-5)
    assert np.abs(res[lc.F32][2] - res[lc.F64][2]).max() <= 1e-5
    assert np.allclose(res[lc.F32][1], res[lc.F64][1], rtol=1e-5, atol=1e-2)
