"""Config 3 cross-check (GPU box): a whole learnVDP fit on a 1 M-row subsample of the benchmark mixture with the fp32
engine (tensor-core tier) and with the fp64 engine: same K, same number of VB iterations, F within 1e-5."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import libcluster_b200 as lc  # noqa: E402

N, D, K = 1_000_000, 128, 16
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
mu, L, w = bench.mixture_params(D, K)
mu_t = torch.tensor(mu, dtype=torch.float32, device=dev)
L_t = torch.tensor(L, dtype=torch.float32, device=dev)
w_t = torch.tensor(w, dtype=torch.float32, device=dev)
X = torch.cat([bench.gen_chunk_torch(torch, dev, c, bench.CHUNK, D, K, mu_t, L_t, w_t)[0] for c in range(N // bench.CHUNK)])
torch.cuda.synchronize()
out = {}
for name, prec in (("f32", lc.F32), ("f64", lc.F64)):
    eng = lc.Engine(0, prec)
    eng.set_data_device(X.data_ptr(), N, D, D)
    t0 = time.perf_counter()
    F = eng.learn(lc.VDP, maxclusters=K)
    out[name] = {"K": int(eng.K), "F": F, "vb_iterations": int(len(eng.trace()[0])), "seconds": time.perf_counter() - t0}
    eng.close()
out["rel_dF"] = abs(out["f32"]["F"] - out["f64"]["F"]) / abs(out["f64"]["F"])
out["same_K"] = out["f32"]["K"] == out["f64"]["K"]
out["same_iterations"] = out["f32"]["vb_iterations"] == out["f64"]["vb_iterations"]
print(json.dumps(out), flush=True)
