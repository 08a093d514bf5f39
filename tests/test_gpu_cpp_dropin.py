"""GPU (-m gpu): the C++ drop-in boundary.  A user program written against the reference's own API
(tests/cpp/dropin_main.cpp, cf. test/cluster_test.cpp) is compiled with this repo's include/libcluster.h
and include/distributions.h (Eigen stand-in: oracle/refshim, real Eigen is not in the image), linked to
liblcb200.so, run on the GPU, and compared with the golden fixtures / the oracle."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def test_cpp_user_program_runs_on_the_engine(tmp_path, testdata):
    X, _ = testdata
    exe = tmp_path / "dropin"
    libdir = os.path.join(ROOT, "libcluster_b200", "_lib")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-std=c++11", "-O1", "-w", "-I" + os.path.join(ROOT, "oracle", "refshim"),
                           "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "dropin_main.cpp"),
                           "-o", str(exe), "-L" + libdir, "-llcb200", "-Wl,-rpath," + libdir])
    data = tmp_path / "x.bin"
    with open(data, "wb") as f:
        f.write(struct.pack("ii", len(X), X[0].shape[1]))
        for g in X:
            f.write(struct.pack("i", g.shape[0]))
            f.write(np.ascontiguousarray(g, dtype=np.float64).tobytes())
    out = subprocess.check_output([str(exe), str(data)], text=True)
    g1, g2 = golden("gmc_groups"), golden("bgmm_xcat")
    m = re.search(r"GMC F (\S+) K (\d+)", out)
    assert m and int(m.group(2)) == int(g1["K"]) and float(m.group(1)) == pytest.approx(float(g1["F"]), rel=1e-5)
    means = np.array([[float(v) for v in ln.split()[1:]] for ln in out.splitlines() if ln.startswith("mean ")])
    assert np.allclose(means, g1["means"], atol=1e-4)
    w0 = np.array([float(v) for v in next(ln for ln in out.splitlines() if ln.startswith("w0 ")).split()[1:]])
    assert np.allclose(w0, np.exp(g1["Elogweight"][0]), atol=1e-4)
    m = re.search(r"BGMM F (\S+) K (\d+) rowsum (\S+)", out)
    assert m and int(m.group(2)) == 3 and float(m.group(1)) == pytest.approx(float(g2["F"]), rel=1e-5)
    assert float(m.group(3)) == pytest.approx(120.0, abs=1e-3)
    Xcat = np.concatenate(list(X), 0)
    # a caller's Dirichlet(5.0) / StickBreak(3.0) prior reaches the fit (src/cluster.cpp:653,684)
    for tag, model, wp in (("BGMM5", po.BGMM, 5.0), ("VDP3", po.VDP, 3.0)):
        mo = po.Model(model, [Xcat])
        Fo = mo.learn(weight_prior=wp)
        m = re.search(tag + r" F (\S+) K (\d+) Elogw (.*)", out)
        assert m, out
        assert int(m.group(2)) == mo.K
        assert float(m.group(1)) == pytest.approx(Fo, rel=1e-5)
        elw = np.array([float(v) for v in m.group(3).split()])
        assert np.allclose(elw, mo.weights(0)[0], atol=1e-4)
        # and the prior matters on this data: the default-prior fit has a different F
        assert abs(Fo - po.Model(model, [Xcat]).learn()) > 1e-4 * abs(Fo)
    c = po.Cluster(po.C_GAUSSWISH, 1.0, 2)
    c.addobs(np.ones(120), Xcat)
    c.update()
    m = re.search(r"OPS N (\S+) Esum (\S+) F (\S+)", out)
    assert float(m.group(1)) == pytest.approx(120.0)
    assert float(m.group(2)) == pytest.approx(c.Eloglike(Xcat).sum(), rel=1e-5)
    assert float(m.group(3)) == pytest.approx(c.fenergy(), rel=1e-6)
    assert "ERR invalid_argument Must specify at least one thread" in out
