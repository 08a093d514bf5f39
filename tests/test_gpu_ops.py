"""GPU (-m gpu): operator surface of include/distributions.h on host buffers --
GaussWish / NormGamma addobs, Eloglike, splitobs -- against the oracle."""
import numpy as np
import pytest

import libcluster_b200 as lc
from conftest import make_blobs
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec,tol", [(lc.F64, 1e-10), (lc.F32, 2e-6)])
@pytest.mark.parametrize("kind,cls", [(po.C_GAUSSWISH, lc.GaussWish), (po.C_NORMGAMMA, lc.NormGamma)])
@pytest.mark.parametrize("N,D", [(1000, 2), (777, 16), (300, 128), (64, 1)])
def test_addobs_update_eloglike_splitobs(prec, tol, kind, cls, N, D):
    X, _ = make_blobs(N, D, 1, seed=N + D, spread=5.0, diag=(kind == po.C_NORMGAMMA))
    rng = np.random.default_rng(D)
    q = rng.uniform(size=N)
    o = po.Cluster(kind, 1.3, D)
    c = cls(1.3, D, precision=prec)
    for part in (slice(0, N // 2), slice(N // 2, N)):     # addobs accumulates (distributions.cpp:310-312)
        o.addobs(q[part], X[part])
        c.addobs(q[part], X[part])
    Ns, xs, xxs = c.get_stats()
    so = o.state()
    assert Ns == pytest.approx(so["N_s"], rel=max(tol, 1e-12))
    assert np.allclose(xs, so["x_s"], rtol=0, atol=tol * (1 + np.abs(so["x_s"]).max()) * 10)
    assert np.allclose(xxs, so["xx_s"], rtol=0, atol=tol * (1 + np.abs(so["xx_s"]).max()) * 10)
    o.update(); c.update()
    assert c.getN() == pytest.approx(o.getN(), rel=max(1e-9, tol))
    assert c.fenergy() == pytest.approx(o.fenergy(), rel=max(10 * tol, 1e-10))
    E, Eo = c.Eloglike(X), o.Eloglike(X)
    assert np.allclose(E, Eo, rtol=10 * tol, atol=10 * tol * (1 + np.abs(Eo).max()))
    s, so_ = c.splitobs(X), o.splitobs(X)
    # points within rounding distance of the hyperplane may fall either side in fp32
    assert (s != so_).mean() <= (0.0 if prec == lc.F64 else 0.002)
    # column-major (Eigen default) input gives the same answer
    assert np.allclose(c.Eloglike(np.asfortranarray(X)), E, rtol=1e-12, atol=1e-12)


def test_operator_argument_checks():
    c = lc.GaussWish(1.0, 3)
    X = np.zeros((10, 2))
    with pytest.raises(lc.InvalidArgument, match="Mismatched dims"):
        c.addobs(np.ones(10), X)                       # distributions.cpp:303-304
    with pytest.raises(lc.InvalidArgument, match="not the same length"):
        c.addobs(np.ones(9), np.zeros((10, 3)))        # :305-306


def test_upload_lanes_give_identical_resident_data():
    """lcb_set_data on a page-locked matrix splits the rows between host-converted fp32 blocks and raw fp64 blocks
    converted on the device (engine.cu: upload_rows_f32); a pageable matrix takes the first lane only.  Both must
    leave the same X on the device, so one VB iteration gives the same F and qZ."""
    import torch

    for N, D, K in ((150_000, 128, 4), (400_000, 6, 3)):   # D = 6: rows padded to 8 columns on the device
        X, z = make_blobs(N, D, K, seed=5, spread=4.0)
        q0 = np.full((N, K), 0.01)
        q0[np.arange(N), z] = 1.0 - 0.01 * (K - 1)
        Xp = torch.empty(N, D, dtype=torch.float64, pin_memory=True)
        Xp.copy_(torch.from_numpy(X))
        out = []
        for src in (X, Xp.numpy()):
            eng = lc.Engine(0, lc.F32)
            eng.set_data(src)
            eng.model_init(lc.BGMM)
            eng.set_qz(q0)
            F, _ = eng.vbem(maxit=0)
            out.append((F, eng.qZ(0)))
            eng.set_data(src)            # a second upload of the same shape reuses the device buffers
            eng.model_init(lc.BGMM)
            eng.set_qz(q0)
            F2, _ = eng.vbem(maxit=0)
            assert F2 == pytest.approx(F, rel=1e-8)
            eng.close()
        # (the statistics are summed with fp64 atomics, and fp32 partial sums depend on the order of arrival: runs agree to ~1e-10, not bit for bit)
        assert out[0][0] == pytest.approx(out[1][0], rel=1e-8)
        assert np.abs(out[0][1] - out[1][1]).max() <= 1e-6
