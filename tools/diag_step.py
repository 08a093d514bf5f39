"""Diagnostic (GPU box): VB iterations with a synchronisation after every launch (LCB_DEBUG_SYNC=1) so that a failing
kernel is named.  usage: diag_step.py D K N [steps]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["LCB_DEBUG_SYNC"] = "1"
import libcluster_b200 as lc  # noqa: E402
from conftest import make_blobs  # noqa: E402

D, K, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
X, z = make_blobs(N, D, K, seed=1, spread=4.0)
q0 = np.zeros((N, K))
q0[np.arange(N), z] = 1.0
eng = lc.Engine(0, lc.F32)
eng.set_data(X)
eng.model_init(lc.BGMM)
eng.set_qz(q0)
try:
    for i in range(steps):
        F = eng.vbem_step()
        print("D", D, "K", K, "N", N, "step", i, "F", F, eng.estep_detail(), flush=True)
except Exception as ex:  # noqa: BLE001
    print("D", D, "K", K, "N", N, "FAILED:", ex, flush=True)
