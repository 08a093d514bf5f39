"""Timeline of the accumulator hand-off in estep_coarse_tc128_kernel (diagnostic build, tools/build_trace.sh).

    LCB_LIB_PATH=libcluster_b200/_lib/trace/liblcb200_trace.so python tools/coarse_trace.py

Prints, in SM clocks and averaged over the items CTA 0 recorded, the legs of one accumulator cycle:
release(i-2) -> issuer sees it -> MMAs issued -> epilogue woken -> loads complete -> release(i)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libcluster_b200 as lc  # noqa: E402
from libcluster_b200 import _native as nat  # noqa: E402


def main(N=1_000_000, K=64, D=128):
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    mu = torch.rand(K, D, device=dev, generator=g) * 20 - 10
    z = torch.randint(0, K, (N,), device=dev, generator=g).to(torch.int32)
    X = mu[z.long()] + torch.randn(N, D, device=dev, generator=g)
    eng = lc.Engine(0, lc.F32)
    eng.set_data_device(X.data_ptr(), N, D, D)
    eng.model_init(lc.BGMM)
    eng.set_labels_device(z.data_ptr(), K)
    L = nat.lib()
    buf = (C.c_ulonglong * (1 << 16))()
    for _ in range(2):
        eng.vbem_step()
        n = L.lcb_debug_read_trace(buf, 1 << 16)
    a = np.frombuffer(buf, dtype=np.uint64, count=n)
    tag = (a >> np.uint64(60)).astype(int)
    who = ((a >> np.uint64(56)) & np.uint64(15)).astype(int)
    item = ((a >> np.uint64(32)) & np.uint64(0xFFFFFF)).astype(int)
    clk = (a & np.uint64(0xFFFFFFFF)).astype(np.int64)
    print("entries", n, "items", item.max() + 1 if n else 0)
    ev = {}
    for t, w, i, c in zip(tag, who, item, clk):
        ev[(t, i)] = c
    items = sorted(i for (t, i) in ev if t == 5)
    legs = {"release(i-2)->seen": [], "seen->issued": [], "issued->woken": [], "woken->loaded": [], "loaded->release": [],
            "cycle release(i-2)->release(i)": [], "per item (release(i-1)->release(i))": []}

    def d(x, y):
        return int((y - x) & 0xFFFFFFFF)

    for i in items:
        if i < 8 or not all((t, i) in ev for t in (1, 2, 3, 4, 5)) or (5, i - 2) not in ev or (5, i - 1) not in ev:
            continue
        legs["release(i-2)->seen"].append(d(ev[(5, i - 2)], ev[(1, i)]))
        legs["seen->issued"].append(d(ev[(1, i)], ev[(2, i)]))
        legs["issued->woken"].append(d(ev[(2, i)], ev[(3, i)]))
        legs["woken->loaded"].append(d(ev[(3, i)], ev[(4, i)]))
        legs["loaded->release"].append(d(ev[(4, i)], ev[(5, i)]))
        legs["cycle release(i-2)->release(i)"].append(d(ev[(5, i - 2)], ev[(5, i)]))
        legs["per item (release(i-1)->release(i))"].append(d(ev[(5, i - 1)], ev[(5, i)]))
    for k, v in legs.items():
        v = np.array([x for x in v if x < 1 << 20])
        if len(v):
            print("%-40s n=%d mean %.0f median %.0f p10 %.0f p90 %.0f" % (k, len(v), v.mean(), np.median(v), np.percentile(v, 10), np.percentile(v, 90)))
    eng.close()


if __name__ == "__main__":
    main()
