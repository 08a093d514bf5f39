"""CPU: pins the oracle against the REFERENCE'S OWN CODE.

oracle/_ref/libcluster_ref.so is built by `make -C oracle ref` from /root/reference/src/*.cpp where they lie,
against minimal stand-ins for the absent Eigen/Boost headers (oracle/refshim).  Wherever that library exists
(the build container; it also travels to the GPU box), the C restatement must reproduce it."""
import numpy as np
import pytest

from conftest import golden, make_blobs, soft_labels
from oracle import pyoracle as po
from oracle import pyref

pytestmark = pytest.mark.skipif(not (pyref.available() or pyref.build()),
                                reason="oracle/_ref not built (no /root/reference here)")

CASES = [("bgmm_xcat", po.BGMM, False), ("vdp_xcat", po.VDP, False), ("dgmm_xcat", po.DGMM, False),
         ("gmc_groups", po.GMC, True), ("sgmc_groups", po.SGMC, True), ("dgmc_groups", po.DGMC, True)]


@pytest.mark.parametrize("name,model,grouped", CASES)
def test_reference_learn_on_its_own_fixture_equals_oracle_and_golden(testdata, name, model, grouped):
    X, _ = testdata
    groups = list(X) if grouped else [np.concatenate(list(X), 0)]
    r = pyref.learn(model, groups)
    g = golden(name)
    assert r.K == int(g["K"])
    assert r.F == pytest.approx(float(g["F"]), rel=1e-12)
    assert np.abs(np.concatenate(r.qZ, 0) - g["qZ"]).max() < 1e-12
    m = po.Model(model, groups)
    m.learn()
    for k in range(r.K):
        c = m.cluster(k)
        assert np.allclose(r.means[k], c["m"], rtol=1e-12, atol=1e-12)
        assert r.cfen[k] == pytest.approx(c["fenergy"], rel=1e-11)
        assert r.N[k] == pytest.approx(c["N"], rel=1e-12)
    for j in range(len(groups)):
        assert np.allclose(r.Elogweight[j], m.weights(j)[0], rtol=1e-12, atol=1e-13)
        assert r.wfen[j] == pytest.approx(m.weights_fenergy(j), rel=1e-11, abs=1e-11)


@pytest.mark.parametrize("model,D,K,diag", [(po.BGMM, 5, 4, False), (po.VDP, 12, 3, False), (po.DGMM, 9, 5, True),
                                            (po.GMC, 3, 3, False), (po.DGMC, 4, 3, True),
                                            # the dimensions of the tensor-core tier (configs 2-4 of BASELINE.json)
                                            (po.BGMM, 64, 4, False), (po.GMC, 64, 3, False), (po.VDP, 128, 3, False)])
def test_reference_vbem_iterations_equal_oracle(model, D, K, diag):
    """vbem<W,C>() of src/cluster.cpp:177 called directly, for every iteration count up to 4."""
    X, z = make_blobs(600, D, K, seed=D * 7, spread=4.0, diag=diag)
    q0 = soft_labels(z, K, seed=D)
    groups = [X[:250], X[250:]] if model in (po.GMC, po.DGMC) else [X]
    for maxit in (0, 1, 3):
        r = pyref.vbem(model, groups, q0, maxit=maxit)
        m = po.Model(model, groups)
        F = m.vbem(q0, maxit=maxit)
        assert r.F == pytest.approx(F, rel=1e-12)
        assert np.abs(np.concatenate(r.qZ, 0) - m.qZ()).max() < 1e-11


def test_reference_learn_with_splits_and_sparse_equals_oracle():
    X, z = make_blobs(1200, 3, 4, seed=21, spread=8.0)
    for model, groups, sparse in [(po.VDP, [X], False), (po.BGMM, [X], False),
                                  (po.GMC, [X[:500], X[500:520], X[520:]], False),
                                  (po.GMC, [X[:500], X[500:520], X[520:]], True)]:
        r = pyref.learn(model, groups, sparse=sparse)
        m = po.Model(model, groups)
        F = m.learn(sparse=sparse)
        assert r.K == m.K
        assert r.F == pytest.approx(F, rel=1e-11)
        assert np.abs(np.concatenate(r.qZ, 0) - m.qZ()).max() < 1e-10


@pytest.mark.parametrize("model,wp", [(po.BGMM, 5.0), (po.VDP, 3.0), (po.DGMM, 0.2), (po.VDP, 0.5)])
def test_reference_keeps_a_callers_weight_prior_like_the_oracle(testdata, model, wp):
    """learnBGMM(X, qZ, Dirichlet(alpha), ...) / learnVDP(X, qZ, StickBreak(c), ...): the fit keeps the caller's prior
    (src/cluster.cpp:653,684) while the split refinements use default-constructed weights (:460-461).  This is what the
    C++ drop-in header forwards to lcb_learn (tests/test_gpu_cpp_dropin.py compares the engine with the oracle)."""
    X, _ = testdata
    Xcat = np.concatenate(list(X), 0)
    r = pyref.learn(model, [Xcat], weight_prior=wp)
    m = po.Model(model, [Xcat])
    F = m.learn(weight_prior=wp)
    assert r.K == m.K
    assert r.F == pytest.approx(F, rel=1e-12)
    assert np.abs(r.qZ[0] - m.qZ()).max() < 1e-11
    assert np.allclose(r.Elogweight[0], m.weights(0)[0], rtol=1e-12, atol=1e-13)
    assert r.wfen[0] == pytest.approx(m.weights_fenergy(0), rel=1e-11, abs=1e-11)
    # and it is not the default-prior fit
    assert abs(F - po.Model(model, [Xcat]).learn()) > 1e-6 * abs(F)
    with pytest.raises(pyref.RefError) as e:
        pyref.learn(model, [Xcat], weight_prior=0.0)     # distributions.cpp:107,234
    assert e.value.code == 1


def test_reference_error_behaviour():
    X, _ = make_blobs(50, 2, 1, seed=1)
    with pytest.raises(pyref.RefError) as e:
        pyref.learn(po.BGMM, [X], nthreads=0)      # cluster.cpp:576-577
    assert e.value.code == 1 and "at least one thread" in str(e.value)
    with pytest.raises(pyref.RefError) as e:
        pyref.learn(po.BGMM, [X], prior=-1.0)      # distributions.cpp:282-283
    assert e.value.code == 1
