"""CPU: the C-ABI library loads, exports every symbol include/libcluster_b200.h
declares, fails loudly without a GPU, and its host-only pieces (weight / cluster
posteriors, packed M-step) agree with the oracle."""
import os
import re

import numpy as np
import pytest

import libcluster_b200 as lc
from conftest import ROOT, make_blobs, soft_labels
from libcluster_b200 import _native as nat
from oracle import pyoracle as po


def test_every_declared_symbol_is_exported():
    hdr = open(os.path.join(ROOT, "include", "libcluster_b200.h")).read()
    declared = set(re.findall(r"\b(lcb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"lcb_allreduce_fn"}
    L = nat.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(nat.SIGNATURES), declared ^ set(nat.SIGNATURES)
    assert b"sm_100a" in L.lcb_version()


def test_operand_packing_paths_agree():
    """tc_pack_cluster picks an F16C path at run time; it must produce the bytes of the portable path."""
    assert nat.lib().lcb_selftest_host_packing() == 0


def test_no_gpu_means_error_not_fallback():
    if nat.lib().lcb_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(lc.CudaError, match="no CPU fallback"):
        lc.Engine()


@pytest.mark.parametrize("kind,cls", [(po.W_DIRICHLET, lc.Dirichlet), (po.W_STICKBREAK, lc.StickBreak),
                                      (po.W_GDIRICHLET, lc.GDirichlet)])
@pytest.mark.parametrize("K", [1, 2, 5, 33])
def test_weight_posteriors(kind, cls, K):
    rng = np.random.default_rng(K)
    Nk = rng.uniform(0, 100, K)
    Nk[rng.integers(K)] = 0.0
    for prior in (None, 0.3, 4.0):
        w = cls(prior)
        o = po.Weight(kind, -1.0 if prior is None else prior)
        w.update(Nk); o.update(Nk)
        assert np.allclose(w.Elogweight(), o.Elogweight(), rtol=1e-12, atol=1e-13)
        assert np.allclose(w.getNk(), Nk)
        assert w.fenergy() == pytest.approx(o.fenergy(), rel=1e-11, abs=1e-11)
    with pytest.raises(lc.InvalidArgument):
        cls(-2.0)
    # C ABI: a negative prior is the "default constructor" sentinel, an explicit 0 is the reference's invalid_argument
    import ctypes as C
    h = C.c_void_p()
    assert nat.lib().lcb_weights_create(C.byref(h), kind, C.c_double(0.0)) == 1  # LCB_EINVAL
    assert b"> 0" in nat.lib().lcb_last_error()


@pytest.mark.parametrize("kind,cls", [(po.C_GAUSSWISH, lc.GaussWish), (po.C_NORMGAMMA, lc.NormGamma)])
@pytest.mark.parametrize("D", [1, 2, 7, 13, 32, 128])
def test_cluster_posteriors_from_stats(kind, cls, D):
    rng = np.random.default_rng(D)
    X = rng.normal(size=(300, D)) * 1.7 - 4
    q = rng.uniform(size=300)
    o = po.Cluster(kind, 0.7, D)
    o.addobs(q, X); o.update()
    st = o.state()
    c = cls(0.7, D)
    c.set_stats(st["N_s"], st["x_s"], st["xx_s"])
    c.update()
    assert c.getN() == pytest.approx(st["N"])
    assert c.getprior() == 0.7
    assert np.allclose(c.getmean(), st["m"], rtol=1e-12)
    cov = st["iW"] / st["nu"] if kind == po.C_GAUSSWISH else st["iW"] * st["nu"]
    assert np.allclose(c.getcov(), cov, rtol=1e-10)
    assert c.fenergy() == pytest.approx(o.fenergy(), rel=1e-10)
    c.clearobs()
    assert c.get_stats()[0] == 0.0
    with pytest.raises(lc.InvalidArgument):
        cls(0.0, D)


def test_non_pd_update_is_domain_error():
    c = lc.GaussWish(1.0, 2)
    c.set_stats(10.0, np.array([1.0, 1.0]), np.array([[-50.0, 0.0], [0.0, -50.0]]))
    with pytest.raises(lc.DomainError):
        c.update()
    d = lc.NormGamma(1.0, 2)
    d.set_stats(10.0, np.array([1.0, 1.0]), np.array([-50.0, -50.0]))
    with pytest.raises(lc.InvalidArgument):
        d.update()


@pytest.mark.parametrize("model,omodel", [(lc.BGMM, po.BGMM), (lc.VDP, po.VDP), (lc.DGMM, po.DGMM), (lc.GMC, po.GMC)])
def test_host_mstep_matches_oracle_iteration(model, omodel):
    """Packed statistics -> posteriors -> parameter free energy == one oracle iteration's M half."""
    X, z = make_blobs(240, 3, 3, seed=2)
    groups = [X[:100], X[100:]] if model == lc.GMC else [X]
    q0 = soft_labels(z, 3, seed=2)
    m = po.Model(omodel, groups)
    m.vbem(q0, maxit=0)
    J, K, D = len(groups), 3, 3
    diag = omodel == po.DGMM
    S = D if diag else D * D
    packed = np.zeros(lc.packed_len(model, J, K, D))
    assert packed.size == J * K + K * (1 + D + S)
    off = 0
    for g in groups:
        packed[off // 1:off + K] = 0
        off += K
    o = 0
    r = 0
    for j, g in enumerate(groups):
        packed[j * K:(j + 1) * K] = q0[r:r + g.shape[0]].sum(0)
        r += g.shape[0]
    for k in range(K):
        c = m.cluster(k)
        b = J * K + k * (1 + D + S)
        packed[b] = c["N_s"]
        packed[b + 1:b + 1 + D] = c["x_s"]
        packed[b + 1 + D:b + 1 + D + S] = c["xx_s"].ravel()
    F, elogw, means, covs = lc.host_mstep(model, packed, J, K, D)
    Fref = sum(m.weights_fenergy(j) for j in range(J)) + sum(m.cluster(k)["fenergy"] for k in range(K))
    assert F == pytest.approx(Fref, rel=1e-10)
    for j in range(J):
        assert np.allclose(elogw[j], m.weights(j)[0], rtol=1e-11, atol=1e-12)
    for k in range(K):
        assert np.allclose(means[k], m.cluster(k)["m"], rtol=1e-11)


def test_shard_rows_partition():
    for N in (0, 1, 7, 100, 12345):
        for world in (1, 2, 3, 8):
            cuts = [lc.shard_rows(N, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == N
            for a, b in zip(cuts[:-1], cuts[1:]):
                assert a[1] == b[0]
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("n", [1, 2, 3, 16, 17, 33, 64, 100, 256, 512])
def test_device_stick_order_is_std_sort(n):
    """The device M step orders sticks with a restatement of libstdc++'s std::sort (mstep_math.hpp): same order as
    the reference's std::sort call (distributions.cpp:146) on random, tie-heavy and adversarial counts."""
    import ctypes as C
    L = nat.lib()
    rng = np.random.default_rng(n)
    cases = [rng.uniform(0, 50, n), np.zeros(n), np.arange(n, dtype=float), np.arange(n, dtype=float)[::-1].copy(),
             rng.integers(0, 3, n).astype(float), np.where(rng.random(n) < 0.7, 0.0, rng.uniform(0, 9, n)),
             np.tile([5.0, 1.0], n)[:n].copy(), np.r_[np.arange(n // 2), np.arange(n - n // 2)].astype(float)]
    # median-of-three killer sequence (drives introsort into its heap-sort fallback for larger n)
    if n >= 4 and n % 2 == 0:
        k = n // 2
        mk = np.zeros(n)
        for i in range(1, k + 1):
            mk[i - 1] = i if i % 2 == 1 else k + i - 1
            mk[k + i - 1] = 2 * i
        cases.append(-mk)
    for v in cases:
        v = np.ascontiguousarray(v, dtype=np.float64)
        order = np.zeros(n, dtype=np.int32)
        diff = L.lcb_selftest_stick_order(v.ctypes.data_as(C.POINTER(C.c_double)), n,
                                          order.ctypes.data_as(C.POINTER(C.c_int)))
        assert diff == 0
        assert sorted(order.tolist()) == list(range(n))
        assert np.all(np.diff(v[order]) <= 0)


def test_dropin_headers_compile_and_link(tmp_path):
    """CPU: the reference-facing C++ headers (include/libcluster.h, include/distributions.h) compile as a user program
    (tests/cpp/dropin_main.cpp, written against the reference's API) against the Eigen stand-in and link to the C-ABI
    library.  Running it needs a GPU (tests/test_gpu_cpp_dropin.py); this only guards the boundary's syntax and
    symbols, including the weight-prior accessor the fits read (src/cluster.cpp:653,684)."""
    import shutil
    import subprocess
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no g++")
    libdir = os.path.join(ROOT, "libcluster_b200", "_lib")
    exe = tmp_path / "dropin"
    subprocess.check_call([cxx, "-std=c++11", "-O0", "-w", "-I" + os.path.join(ROOT, "oracle", "refshim"),
                           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "dropin_main.cpp"),
                           "-o", str(exe), "-L" + libdir, "-llcb200", "-Wl,-rpath," + libdir])
    assert exe.exists()
    hdr = open(os.path.join(ROOT, "include", "libcluster.h")).read()
    assert "common_weight_prior" in hdr and "lcb_learn(g.e, model, clusterprior, wprior" in hdr
