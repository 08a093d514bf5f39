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
