"""CPU: the oracle against its pins -- the derived golden fixtures, the
independent numpy mirror, mpmath special functions, and the reference's own
documented invariants (README.md:285-287 rows of qZ sum to 1; testdata.h:27
three well separated clusters)."""
import numpy as np
import pytest

from conftest import golden, make_blobs, soft_labels
from oracle import np_oracle as npo
from oracle import pyoracle as po

CASES = [("bgmm_xcat", po.BGMM, False), ("vdp_xcat", po.VDP, False), ("dgmm_xcat", po.DGMM, False),
         ("gmc_groups", po.GMC, True), ("sgmc_groups", po.SGMC, True), ("dgmc_groups", po.DGMC, True)]


@pytest.mark.parametrize("name,model,grouped", CASES)
def test_oracle_matches_golden(testdata, name, model, grouped):
    X, _ = testdata
    groups = list(X) if grouped else [np.concatenate(list(X), 0)]
    g = golden(name)
    m = po.Model(model, groups)
    F = m.learn()
    assert m.K == int(g["K"])
    assert F == pytest.approx(float(g["F"]), rel=1e-12)
    Ft, Kt = m.trace()
    assert np.allclose(Ft, g["trace_F"], rtol=1e-12)
    assert np.array_equal(Kt, g["trace_K"])
    q = m.qZ()
    assert np.abs(q - g["qZ"]).max() < 1e-12
    assert np.allclose(q.sum(1), 1.0, atol=1e-9)


def test_designed_answer_config1(testdata):
    """testdata.h:27,38,50 -- clusters near [0,0], [-10,10], [10,10]."""
    X, _ = testdata
    m = po.Model(po.BGMM, [np.concatenate(list(X), 0)])
    m.learn()
    assert m.K == 3
    means = np.stack([m.cluster(k)["m"] for k in range(3)])
    for target in ([0, 0], [-10, 10], [10, 10]):
        assert np.min(np.linalg.norm(means - np.array(target), axis=1)) < 1.0


def test_digamma_against_mpmath():
    mpmath = pytest.importorskip("mpmath")
    for x in [0.5, 1.0, 1.5, 2.0, 3.7, 9.99, 10.0, 64.5, 1e3, 5e7]:
        assert po.digamma(x) == pytest.approx(float(mpmath.digamma(x)), rel=2e-15, abs=2e-15)


@pytest.mark.parametrize("mname,model", [("BGMM", po.BGMM), ("VDP", po.VDP), ("DGMM", po.DGMM)])
def test_c_oracle_equals_numpy_mirror_on_blobs(mname, model):
    X, _ = make_blobs(400, 3, 4, seed=11)
    m = po.Model(model, [X])
    F = m.learn()
    Fn, qn, wn, cn, tr = npo.learn(mname, [X])
    assert m.K == len(cn)
    assert F == pytest.approx(Fn, rel=1e-9)
    assert np.abs(m.qZ() - qn[0]).max() < 1e-8


def test_vbem_from_given_labels_matches_mirror():
    X, z = make_blobs(300, 4, 3, seed=5)
    q0 = soft_labels(z, 3, seed=5)
    m = po.Model(po.BGMM, [X])
    F = m.vbem(q0, maxit=4)
    W, Cc = npo.MODELS["BGMM"]
    q = [q0.copy()]
    w, c, tr = [], [], []
    Fn = npo.vbem([X], q, w, c, W, Cc, 1.0, 4, False, tr)
    assert len(tr) == len(m.trace()[0]) <= 5
    assert F == pytest.approx(Fn, rel=1e-10)
    assert np.abs(m.qZ() - q[0]).max() < 1e-9


def test_operator_level_restatement():
    rng = np.random.default_rng(3)
    X = rng.normal(size=(50, 3)) + 2
    q = rng.uniform(size=50)
    for kind, cls in [(po.C_GAUSSWISH, npo.GaussWish), (po.C_NORMGAMMA, npo.NormGamma)]:
        a = po.Cluster(kind, 1.0, 3)
        b = cls(1.0, 3)
        a.addobs(q, X); a.update()
        b.addobs(q, X); b.update()
        assert np.allclose(a.Eloglike(X), b.Eloglike(X), rtol=1e-11)
        assert a.fenergy() == pytest.approx(b.fenergy(), rel=1e-11)
        assert np.array_equal(a.splitobs(X), b.splitobs(X))
    with pytest.raises(po.OracleError):
        po.Cluster(po.C_GAUSSWISH, -1.0, 2)


def test_free_energy_is_monotone_on_trace(testdata):
    """cluster.cpp:229-230 -- inside each vbem call F never rises by more than FENGYDEL."""
    X, _ = testdata
    m = po.Model(po.VDP, [np.concatenate(list(X), 0)])
    m.learn()
    F, K = m.trace()
    assert np.isfinite(F).all() and len(F) > 10
