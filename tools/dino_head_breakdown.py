"""Per-entry-point device time of one fused DINO head step, eager launches, CUDA events around
every C-ABI call (median of `reps` steps).    python tools/dino_head_breakdown.py [B ncrops K D]"""
import json
import os
import sys
from collections import defaultdict

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lafs_cvpr2024_b200 as P  # noqa: E402
from lafs_cvpr2024_b200 import _lib  # noqa: E402

if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]] + [256, 6, 65536, 256][len(sys.argv) - 1:]
    B, ncrops, K, D = a[:4]
    dev = torch.device("cuda", 0)
    torch.manual_seed(11)
    xs = torch.randn(ncrops * B, D, device=dev)
    xt = torch.randn(2 * B, D, device=dev)
    vs = torch.randn(K, D, device=dev) * 0.02
    vt = vs + torch.randn(K, D, device=dev) * 0.002
    one = torch.ones(K, device=dev)
    center = torch.randn(K, device=dev) * 0.05
    gout = torch.ones((), device=dev)

    def step():
        loss, colsum, saved = P.dino_head_forward(xs, xt, vs, one, vt, one, center, ncrops, 10.0, 25.0)
        return P.dino_head_backward(saved, gout)

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    real_call = _lib.call
    log = []

    def timed_call(name, *args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        real_call(name, *args)
        e1.record()
        log.append((name, e0, e1))

    import lafs_cvpr2024_b200.dino_head as DH
    DH._lib.call = timed_call
    reps = 10
    per = defaultdict(list)
    order = []
    for r in range(reps):
        log.clear()
        step()
        torch.cuda.synchronize()
        seen = defaultdict(int)
        for name, e0, e1 in log:
            key = "%s#%d" % (name, seen[name])
            seen[name] += 1
            per[key].append(e0.elapsed_time(e1))
            if r == 0:
                order.append(key)
    out = {k: round(float(np.median(per[k])) * 1e3, 1) for k in order}
    out["sum_us"] = round(sum(out.values()), 1)
    print(json.dumps(out))
