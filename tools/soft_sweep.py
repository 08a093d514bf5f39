"""Soft-regime sweep (GPU box): candidate pairs per row and throughput against the spread of the cluster means."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import libcluster_b200 as lc  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
for sp in [float(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0.7,0.6,0.45,0.4,0.3").split(",")]:
    print(json.dumps(bench.soft_point(torch, lc, dev, 128, 64, rows, sp)), flush=True)
