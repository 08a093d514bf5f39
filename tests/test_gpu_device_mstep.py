"""GPU (-m gpu): the device-resident M step (mstep.cu, engine_dev.cu) against the oracle and against the host M step
of the same engine (LCB_HOST_MSTEP=1): posteriors to 1e-10, F to 1e-9 (fp64 engine), and the per-iteration
synchronisation budget of the default path."""
import os

import numpy as np
import pytest

import libcluster_b200 as lc
from conftest import make_blobs, soft_labels
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _engine(prec, host_mstep):
    if host_mstep:
        os.environ["LCB_HOST_MSTEP"] = "1"
    else:
        os.environ.pop("LCB_HOST_MSTEP", None)
    try:
        return lc.Engine(0, prec)
    finally:
        os.environ.pop("LCB_HOST_MSTEP", None)


def _posteriors(eng):
    out = []
    for k in range(eng.K):
        c = eng.cluster(k)
        out.append((c["N"], c["mean"], c["cov"], c["fenergy"]))
    return out


@pytest.mark.parametrize("model,D,K", [(lc.BGMM, 2, 3), (lc.VDP, 7, 5), (lc.VDP, 64, 17), (lc.BGMM, 128, 9),
                                       (lc.DGMM, 33, 6), (lc.VDP, 160, 4), (lc.VDP, 200, 3)])
def test_device_mstep_matches_oracle_fp64(model, D, K):
    N = 3000
    X, z = make_blobs(N, D, K, seed=D + K, spread=4.0, diag=model == lc.DGMM)
    q0 = soft_labels(z, K, seed=K)
    m = po.Model(model, [X])
    m.vbem(q0, maxit=3)
    Fo, _ = m.trace()
    res = {}
    for host in (False, True):
        eng = _engine(lc.F64, host)
        eng.set_data(X)
        eng.model_init(model)
        eng.set_qz(q0)
        eng.vbem(maxit=3)
        res[host] = (np.array(eng.trace()[0]), eng.qZ(0), _posteriors(eng), eng.group_weights(0))
        eng.close()
    Fd, qd, pd, wd = res[False]
    Fh, qh, ph, wh = res[True]
    assert len(Fd) == len(Fo)
    assert np.allclose(Fd, Fo, rtol=1e-9), (Fd, Fo)
    assert np.allclose(Fd, Fh, rtol=1e-10), (Fd, Fh)
    assert np.abs(qd - m.qZ()).max() <= 1e-8
    for (Nd, md, cd, fd), (Nh, mh, ch, fh) in zip(pd, ph):
        assert Nd == pytest.approx(Nh, rel=1e-10, abs=1e-10)
        assert np.allclose(md, mh, rtol=1e-10, atol=1e-10)
        assert np.allclose(cd, ch, rtol=1e-10, atol=1e-10)
        assert fd == pytest.approx(fh, rel=1e-10, abs=1e-9)
    assert np.allclose(wd[1], wh[1], rtol=1e-10, atol=1e-12)
    assert wd[2] == pytest.approx(wh[2], rel=1e-10, abs=1e-10)
    # posteriors against the oracle's (getcov() = iW / nu for GaussWish, L * nu for NormGamma)
    for k in range(K):
        co = m.cluster(k)
        assert np.allclose(pd[k][1], co["m"], rtol=1e-9, atol=1e-9)
        cov_o = co["iW"] / co["nu"] if model != lc.DGMM else co["iW"] * co["nu"]
        assert np.allclose(pd[k][2], cov_o, rtol=1e-8, atol=1e-9)
        assert pd[k][3] == pytest.approx(co["fenergy"], rel=1e-9, abs=1e-8)


@pytest.mark.parametrize("model", [lc.GMC, lc.SGMC, lc.DGMC])
@pytest.mark.parametrize("sparse", [False, True])
def test_device_mstep_grouped_models(model, sparse):
    D, K, J = 5, 6, 7
    X, z = make_blobs(2800, D, K, seed=3, spread=5.0, diag=model == lc.DGMC)
    groups = np.array_split(np.arange(X.shape[0]), J)
    Xs = [X[g] for g in groups]
    q0 = soft_labels(z, K, seed=2)
    res = {}
    for host in (False, True):
        eng = _engine(lc.F64, host)
        eng.set_data(Xs)
        eng.model_init(model, sparse=sparse)
        eng.set_qz(q0)
        eng.vbem(maxit=4)
        res[host] = (np.array(eng.trace()[0]), np.concatenate([eng.qZ(j) for j in range(J)]),
                     [eng.group_weights(j)[1] for j in range(J)])
        eng.close()
    assert np.allclose(res[False][0], res[True][0], rtol=1e-10)
    assert np.abs(res[False][1] - res[True][1]).max() <= 1e-9
    for a, b in zip(res[False][2], res[True][2]):
        assert np.allclose(a, b, rtol=1e-10, atol=1e-12)


def test_steady_state_step_budget():
    """One vbem_step of the default path: a single host synchronisation, no collective on one GPU, and the
    tensor-core kernels on the D = 128 path."""
    D, K, N = 128, 16, 20000
    X, z = make_blobs(N, D, K, seed=5, spread=6.0)
    eng = lc.Engine(0, lc.F32)
    eng.set_data(X)
    eng.model_init(lc.BGMM)
    q0 = np.zeros((N, K))
    q0[np.arange(N), z] = 1.0
    eng.set_qz(q0)
    F = [eng.vbem_step() for _ in range(4)]
    c = eng.step_counts()
    assert c["device_mstep"]
    assert c["host_syncs"] <= 1, c
    assert c["collectives"] == 0, c
    assert eng.estep_detail()["path"] == 1
    assert all(F[i + 1] <= F[i] + 1e-6 * abs(F[i]) for i in range(3)), F
    # the host objects follow on demand
    m = po.Model(po.BGMM, [X])
    m.vbem(q0, maxit=3)
    Fo = m.trace()[0]            # the oracle stops when it has converged (cluster.cpp:235), vbem_step never does
    assert np.allclose(F[:len(Fo)], Fo, rtol=1e-5)
    assert F[3] == pytest.approx(Fo[-1], rel=1e-5)
    assert np.abs(eng.qZ(0) - m.qZ()).max() <= 1e-5
    for k in range(K):
        assert np.allclose(eng.cluster(k)["mean"], m.cluster(k)["m"], rtol=1e-5, atol=1e-5)
    eng.close()
