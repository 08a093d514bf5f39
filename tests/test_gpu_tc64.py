"""GPU (-m gpu): the tcgen05 tier on 64-dimensional rows (BASELINE configs 2 and 4: D = 64).

The D = 128 kernels run with DIM = 64: the operand blobs keep R_k in the upper-left 64 x 64 corner, K block 1 and the
upper accumulator columns are skipped (tc_kernels.cu).  Checked against the oracle (F of every iteration, qZ, N_k)
at D = 64, K = 32 (config 2's shape), against the SIMT tier and the fp64 engine, for the dense kernel (K < 8), the
two-level pass, grouped models with list reuse, and a full fit with splits."""
import os

import numpy as np
import pytest

import libcluster_b200 as lc
from conftest import make_blobs, soft_labels
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

D = 64


def _engine(env=None, prec=lc.F32):
    env = env or {}
    for k, v in env.items():
        os.environ[k] = v
    try:
        return lc.Engine(0, prec)
    finally:
        for k in env:
            os.environ.pop(k, None)


def _run(X, q0, model=lc.BGMM, maxit=2, env=None, prec=lc.F32, prior=1.0, sparse=False):
    eng = _engine(env, prec)
    eng.set_data(X)
    eng.model_init(model, prior=prior, sparse=sparse)
    eng.set_qz(q0)
    eng.vbem(maxit=maxit)
    q = eng.qZ()
    q = np.concatenate(q, 0) if isinstance(X, list) else q[0]
    out = (np.array(eng.trace()[0]), q, eng.estep_detail(), eng.group_weights(0)[0])
    eng.close()
    return out


@pytest.mark.parametrize("N,K,spread,prior", [(6000, 32, 3.0, 1.0), (4100, 32, 3.0, 10.0), (3000, 5, 2.0, 1.0),
                                              (5000, 17, 1.0, 1.0)])
def test_tc64_matches_oracle(N, K, spread, prior):
    X, z = make_blobs(N, D, K, seed=N + K, spread=spread)
    q0 = soft_labels(z, K, seed=K, noise=0.2)
    m = po.Model(po.BGMM, [X])
    m.vbem(q0, prior=prior, maxit=2)
    Fo = np.array(m.trace()[0])
    F, q, det, Nk = _run(X, q0, prior=prior)
    # path 1: two-level pass; 2 / 0: it found too many candidates and the dense tensor-core kernel ran (this and the
    # following iterations); K < 8 always runs the dense kernel
    assert det["path"] in ((0, 1, 2) if K >= 8 else (0,)), det
    assert len(F) == len(Fo) and np.allclose(F, Fo, rtol=1e-5, atol=0), (F, Fo)
    # three iterations: 1e-5 where the clusters are separated; the two soft cases (broad prior / spread 1: many
    # candidate pairs per row) sit just below 1e-5 and are held to 2e-5
    assert np.abs(q - m.qZ()).max() <= (1e-5 if (prior == 1.0 and spread >= 2.0) else 2e-5)
    assert np.allclose(Nk, m.weights(0)[1], rtol=1e-5, atol=1e-3)
    # the SIMT tier of the same engine and the fp64 engine agree as well
    Fs, qs, dets, _ = _run(X, q0, prior=prior, env={"LCB_DISABLE_TC": "1"})
    assert np.allclose(F, Fs, rtol=1e-5) and np.abs(q - qs).max() <= 2e-5   # each is within 1e-5 of the oracle
    F64, q64, _, _ = _run(X, q0, prior=prior, prec=lc.F64)
    assert np.allclose(F, F64, rtol=1e-5) and np.abs(q - q64).max() <= 1e-5


def test_tc64_two_level_levels_and_dense_agree():
    N, K = 9000, 32
    X, z = make_blobs(N, D, K, seed=77, spread=2.0)
    q0 = soft_labels(z, K, seed=3)
    Fd, qd, detd, _ = _run(X, q0, maxit=1, env={"LCB_TC_TWO_LEVEL": "0"})
    Ft, qt, dett, _ = _run(X, q0, maxit=1)
    assert detd["path"] == 0 and dett["path"] == 1, (detd, dett)
    assert dett["pairs"] >= N
    assert np.abs(qt - qd).max() <= 2e-6
    assert np.allclose(Ft, Fd, rtol=1e-7)
    assert np.allclose(qt.sum(1), 1.0, atol=1e-5)


@pytest.mark.parametrize("model,omodel,sparse", [(lc.GMC, po.GMC, False), (lc.SGMC, po.SGMC, True)])
def test_tc64_grouped_models_with_list_reuse(model, omodel, sparse):
    """Grouped models at D = 64: per-group weights in the E pass, N_jk from the candidate lists of the previous
    E pass (gather_list_q with group ids) from the second iteration on."""
    K, J = 12, 5
    X, z = make_blobs(7000, D, K, seed=9, spread=6.0)
    cuts = [0, 900, 2500, 2500 + 1, 5200, 7000]
    groups = [X[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    q0 = soft_labels(z, K, seed=5, noise=0.02)   # nearly hard labels: the first M step keeps the clusters apart
    m = po.Model(omodel, groups)
    m.vbem(q0, maxit=3, sparse=sparse)
    Fo = np.array(m.trace()[0])
    F, q, det, _ = _run(groups, q0, model=model, maxit=3, sparse=sparse)
    assert det["path"] == 1, det
    if not sparse:
        assert len(F) == len(Fo) and np.allclose(F, Fo, rtol=1e-5)
    assert np.abs(q - m.qZ()).max() <= 1e-5
    eng = _engine()
    eng.set_data(groups); eng.model_init(model, sparse=sparse); eng.set_qz(q0); eng.vbem(maxit=3)
    for j in range(J):
        assert np.allclose(eng.group_weights(j)[0], m.weights(j)[1], rtol=1e-5, atol=2e-3)
    eng.close()


def test_tc64_full_fit_with_splits():
    X, _ = make_blobs(5000, D, 6, seed=31, spread=4.0)
    m = po.Model(po.VDP, [X])
    Fo = m.learn()
    eng = _engine()
    eng.set_data(X)
    F = eng.learn(lc.VDP)
    assert eng.K == m.K
    assert F == pytest.approx(Fo, rel=1e-5)
    assert np.abs(eng.qZ(0) - m.qZ()).max() <= 1e-5
    eng.close()


def test_tc64_size_independent_properties_at_scale():
    """Config 2's shape at a size the oracle cannot reach (N = 1M, D = 64, K = 32): the tensor-core tier against the
    fp64 engine over three iterations -- F to 1e-5, qZ to 1e-5, rows of qZ sum to one, sum_k N_k = N."""
    import torch
    N, K = 1 << 20, 32
    gen = torch.Generator(device="cuda").manual_seed(4321)
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
        if prec == lc.F32:
            assert eng.estep_detail()["path"] == 1
        Nk = eng.group_weights(0)[0]
        q = eng.qZ(0)
        res[prec] = (Fs, Nk, q)
        assert np.allclose(q.sum(1), 1.0, atol=2e-6)
        assert Nk.sum() == pytest.approx(N, rel=1e-9)
        eng.close()
    assert np.allclose(res[lc.F32][0], res[lc.F64][0], rtol=1e-5)
    assert np.abs(res[lc.F32][2] - res[lc.F64][2]).max() <= 1e-5
