"""Diagnostic: learnBGMM on the reference fixture with progress marks (run on the GPU box)."""
import faulthandler
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(40, exit=False)
import libcluster_b200 as lc  # noqa: E402

d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "testdata.npz"))
X = np.concatenate(list(d["X"]), 0)
prec = lc.F64 if "f64" in sys.argv else lc.F32
eng = lc.Engine(0, prec)
eng.set_data(X)
print("data set", flush=True)
if "vbem" in sys.argv:
    eng.model_init(lc.BGMM)
    eng.set_qz(np.ones((X.shape[0], 1)))
    t = time.time()
    print("vbem K=1:", eng.vbem(maxit=-1), time.time() - t, flush=True)
t = time.time()
print("learn:", eng.learn(lc.BGMM, verbose=True), time.time() - t, flush=True)
print("trace", eng.trace(), flush=True)
