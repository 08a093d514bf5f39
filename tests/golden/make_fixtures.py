"""Generates tests/golden/*.npz.  Run from the repo root in the build container:

    python tests/golden/make_fixtures.py

1. testdata.npz  -- the INPUT fixtures of the reference's own smoke tests
   (makeXdata / makeOdata literals, /root/reference/test/testdata.h:30-220),
   parsed from the header where it lies (needs /root/reference).
2. golden_*.npz  -- OUTPUTS for those inputs.  The reference holds no golden
   numbers, so these are DERIVED: produced by oracle/vb_oracle.c
   and accepted only if oracle/np_oracle.py (independent numpy/scipy mirror)
   agrees to 1e-9, and -- when oracle/_ref/ exists -- the reference's own
   sources compiled against oracle/refshim agree too (see oracle/README.md).
"""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def parse_testdata(path="/root/reference/test/testdata.h"):
    src = open(path).read()
    def blocks(fn):
        body = src[src.index("void " + fn):]
        body = body[:body.index("\n}\n")]
        out = []
        for m in re.finditer(r"<<(.*?);", body, re.S):
            vals = [float(v) for v in re.findall(r"-?\d+\.\d+", m.group(1))]
            out.append(np.array(vals).reshape(-1, 2))
        return out
    return blocks("makeXdata"), blocks("makeOdata")


def main():
    X, O = parse_testdata()
    assert len(X) == 12 and all(x.shape == (10, 2) for x in X)
    assert len(O) == 2 and all(o.shape == (6, 2) for o in O)
    np.savez(os.path.join(OUT, "testdata.npz"), X=np.stack(X), O=np.stack(O))

    from oracle import np_oracle as npo
    from oracle import pyoracle as po
    Xcat = np.concatenate(X, 0)
    cases = {
        "bgmm_xcat": (po.BGMM, "BGMM", [Xcat]),
        "vdp_xcat": (po.VDP, "VDP", [Xcat]),
        "dgmm_xcat": (po.DGMM, "DGMM", [Xcat]),
        "gmc_groups": (po.GMC, "GMC", X),
        "sgmc_groups": (po.SGMC, "SGMC", X),
        "dgmc_groups": (po.DGMC, "DGMC", X),
    }
    for name, (mid, mname, groups) in cases.items():
        m = po.Model(mid, groups)
        F = m.learn(prior=1.0, maxclusters=-1)
        Ft, Kt = m.trace()
        Fn, qn, wn, cn, trn = npo.learn(mname, groups)
        qn = np.concatenate(qn, 0)
        q = m.qZ()
        assert m.K == len(cn), (name, m.K, len(cn))
        assert len(trn) == len(Ft), (name, len(trn), len(Ft))
        assert np.allclose([t[0] for t in trn], Ft, rtol=1e-9, atol=0), name
        assert abs(F - Fn) <= 1e-9 * abs(F), (name, F, Fn)
        assert np.abs(q - qn).max() < 1e-9, name
        from oracle import pyref
        if pyref.available() or pyref.build():      # the reference's own sources (oracle/_ref)
            r = pyref.learn(mid, groups)
            assert r.K == m.K and abs(r.F - F) <= 1e-12 * abs(F), (name, r.F, F)
            assert np.abs(np.concatenate(r.qZ, 0) - q).max() < 1e-12, name
        means = np.stack([m.cluster(k)["m"] for k in range(m.K)])
        covs = np.stack([m.cluster(k)["iW"] / m.cluster(k)["nu"] for k in range(m.K)])
        elogw = np.stack([m.weights(j)[0] for j in range(len(groups))])
        np.savez(os.path.join(OUT, "golden_%s.npz" % name), F=F, K=m.K, qZ=q, trace_F=Ft,
                 trace_K=Kt, means=means, covs=covs, Elogweight=elogw)
        print("%-12s K=%d F=%.10f iters=%d  (C oracle == numpy mirror == reference build)" % (name, m.K, F, len(Ft)))


if __name__ == "__main__":
    main()
