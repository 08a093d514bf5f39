"""Diagnostic (GPU box): where the fp32 engine loses accuracy on two overlapping 128-D clusters
(tests/test_gpu_tc.py::test_tc_overlapping_clusters_*): E pass vs S pass, tensor-core vs SIMT."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import libcluster_b200 as lc  # noqa: E402
from conftest import soft_labels  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

rng = np.random.default_rng(0)
D, N = 128, 6000
base = rng.uniform(-20, 20, size=D)
X = np.concatenate([base + rng.normal(size=(N // 2, D)), base + 0.15 + 1.05 * rng.normal(size=(N // 2, D))])
z = np.repeat([0, 1], N // 2)
q0 = soft_labels(z, 2, seed=1, noise=0.6)
ref = {}
for it in (0, 3):
    m = po.Model(po.VDP, [X])
    m.vbem(q0, maxit=it)
    ref[it] = (m.qZ(), np.array(m.trace()[0]))
modes = [("default", {}), ("simt S pass", {"LCB_TC_SSTAT": "0"}), ("all SIMT", {"LCB_DISABLE_TC": "1"}),
         ("host M step", {"LCB_HOST_MSTEP": "1"})]
for name, env in modes:
    for k, v in env.items():
        os.environ[k] = v
    try:
        for it in (0, 3):
            eng = lc.Engine(0, lc.F32)
            eng.set_data(X)
            eng.model_init(lc.VDP)
            eng.set_qz(q0)
            eng.vbem(maxit=it)
            dq = np.abs(eng.qZ(0) - ref[it][0]).max()
            dF = np.abs(eng.trace()[0] / ref[it][1] - 1).max()
            print("%-12s iterations %d: max|dq| %.3e  max rel dF %.3e" % (name, it + 1, dq, dF), flush=True)
            eng.close()
    finally:
        for k in env:
            os.environ.pop(k, None)
