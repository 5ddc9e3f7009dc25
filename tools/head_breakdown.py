"""Per-kernel timing of the margin-head step on one B200 (development aid): every C entry point of the
step timed with CUDA events, for the kernel variants selected by the LAFS_HEAD_1SM / LAFS_DW_* switches.
    python tools/head_breakdown.py [cfg3|cfg4|cfg3d768]"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lafs_cvpr2024_b200 as P  # noqa: E402
from lafs_cvpr2024_b200 import _lib  # noqa: E402

CFG = {"cfg3": (512, 93431, 512), "cfg4": (1024, 205990, 512), "cfg3d768": (512, 93431, 768)}


def timeit(fn, warmup=3, iters=15):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")   # 256 MB > L2
    for a, b in ev:
        flush.zero_()
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return round(ts[len(ts) // 2] * 1e3, 1)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
    B, C, D = CFG[name]
    torch.manual_seed(0)
    h = P.CosFace(D, C, None).cuda()
    x = torch.randn(B, D, device="cuda")
    lab = torch.randint(0, C, (B,), device="cuda")
    L = _lib.lib()
    st = _lib.stream
    e_hat = torch.empty(B, D, dtype=torch.bfloat16, device="cuda"); inv_e = torch.empty(B, device="cuda")
    w_hat = torch.empty(C, D, dtype=torch.bfloat16, device="cuda"); inv_w = torch.empty(C, device="cuda")
    stats = torch.empty(B, 4, device="cuda")
    nb = L.lafs_head_workspace_bytes(B, C, D)
    ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
    loss = torch.empty((), device="cuda"); lse2 = torch.empty(B, device="cuda"); go = torch.ones((), device="cuda")
    ldg = (C + 7) // 8 * 8
    G = torch.empty(B, ldg, dtype=torch.bfloat16, device="cuda")
    nbb = L.lafs_head_bwd_workspace_bytes(B, C, D)
    wsb = torch.empty(nbb, dtype=torch.uint8, device="cuda")
    de = torch.empty(B, D, device="cuda"); dw = torch.empty(C, D, device="cuda")
    w = h.weight.detach()
    calls = {
        "normalize_w": lambda: _lib.call("lafs_normalize_rows", w.data_ptr(), 0, C, D, w_hat.data_ptr(), inv_w.data_ptr(), st()),
        "normalize_e": lambda: _lib.call("lafs_normalize_rows", x.data_ptr(), 0, B, D, e_hat.data_ptr(), inv_e.data_ptr(), st()),
        "fwd_stats+merge": lambda: _lib.call("lafs_head_fwd", e_hat.data_ptr(), w_hat.data_ptr(), lab.data_ptr(), None, 1.0, B, C, D, 0,
                                             64.0, 0.4, 0, stats.data_ptr(), ws.data_ptr(), nb, st()),
        "loss": lambda: _lib.call("lafs_head_loss", stats.data_ptr(), lab.data_ptr(), None, 1.0, B, lse2.data_ptr(), loss.data_ptr(), st()),
        "grad_logits": lambda: _lib.call("lafs_head_grad_logits", e_hat.data_ptr(), w_hat.data_ptr(), lab.data_ptr(), None, 1.0, B, C, D, 0,
                                         64.0, 0.4, 0, lse2.data_ptr(), go.data_ptr(), 64.0 / B, G.data_ptr(), ldg, st()),
        "bwd_embed(dE)": lambda: _lib.call("lafs_head_bwd_embed", G.data_ptr(), ldg, w_hat.data_ptr(), B, C, D, de.data_ptr(),
                                           wsb.data_ptr(), nbb, st()),
        "bwd_weight(dW+jac)": lambda: _lib.call("lafs_head_bwd_weight", G.data_ptr(), ldg, e_hat.data_ptr(), w_hat.data_ptr(),
                                                inv_w.data_ptr(), B, C, D, dw.data_ptr(), st()),
    }
    res = {"cfg": name, "B": B, "C": C, "D": D,
           "env": {k: os.environ.get(k) for k in ("LAFS_HEAD_1SM", "LAFS_DW_UNFUSED", "LAFS_DW_CLUSTER")}}
    for k, fn in calls.items():
        fn()
    torch.cuda.synchronize()
    for k, fn in calls.items():
        res[k + "_us"] = timeit(fn)
    res["sum_us"] = round(sum(v for k, v in res.items() if k.endswith("_us")), 1)
    fl = 2.0 * B * C * D
    res["fwd_TFLOPs"] = round(fl / res["fwd_stats+merge_us"] / 1e6, 1)
    res["grad_TFLOPs"] = round(fl / res["grad_logits_us"] / 1e6, 1)
    res["dE_TFLOPs"] = round(fl / res["bwd_embed(dE)_us"] / 1e6, 1)
    res["dW_TFLOPs"] = round(fl / res["bwd_weight(dW+jac)_us"] / 1e6, 1)
    xg = x.clone().requires_grad_(True)

    def step():
        xg.grad = None; h.weight.grad = None
        h.forward_loss(xg, lab).backward()
    res["step_eager_us"] = timeit(step, warmup=3, iters=10)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
