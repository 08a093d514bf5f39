"""Diagnostic (GPU box): error of the fp32 engine on the headline shape in the soft regime
(tests/test_gpu_two_level.py::test_headline_shape_against_oracle) per arithmetic path."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import libcluster_b200 as lc  # noqa: E402
from conftest import make_blobs, soft_labels  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

D, K, N = 128, 64, 4096
spread = float(sys.argv[1]) if len(sys.argv) > 1 else 3.5
X, z = make_blobs(N, D, K, seed=5, spread=spread)
q0 = soft_labels(z, K, seed=1, noise=0.2)
ref = {}
for it in (0, 2):
    m = po.Model(po.BGMM, [X])
    m.vbem(q0, prior=10.0, maxit=it)
    ref[it] = (m.qZ(), np.array(m.trace()[0]))
modes = [("default", {}, lc.F32), ("simt S pass", {"LCB_TC_SSTAT": "0"}, lc.F32), ("dense TC E pass", {"LCB_TC_TWO_LEVEL": "0"}, lc.F32),
         ("dense TC + simt S", {"LCB_TC_TWO_LEVEL": "0", "LCB_TC_SSTAT": "0"}, lc.F32),
         ("all SIMT", {"LCB_DISABLE_TC": "1"}, lc.F32), ("fp64", {}, lc.F64)]
for name, env, prec in modes:
    for k, v in env.items():
        os.environ[k] = v
    try:
        for it in (0, 2):
            eng = lc.Engine(0, prec)
            eng.set_data(X)
            eng.model_init(lc.BGMM, prior=10.0)
            eng.set_qz(q0)
            eng.vbem(maxit=it)
            q = eng.qZ(0)
            dq = np.abs(q - ref[it][0])
            i, j = np.unravel_index(dq.argmax(), dq.shape)
            dF = np.abs(eng.trace()[0] / ref[it][1] - 1).max()
            print("%-18s iterations %d: max|dq| %.3e (q there %.4f)  max rel dF %.3e  path %d pairs/row %.2f" % (
                name, it + 1, dq.max(), ref[it][0][i, j], dF, eng.estep_detail()["path"], eng.estep_detail()["pairs"] / N), flush=True)
            eng.close()
    finally:
        for k in env:
            os.environ.pop(k, None)
