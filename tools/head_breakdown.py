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
    from lafs_cvpr2024_b200.margin_head import _round8
    ldg = _round8(C)
    G = torch.empty(B, ldg, dtype=torch.bfloat16, device="cuda")
    ldt = (C + 31) // 32 * 32
    tpart = torch.empty(4 * ((B + 127) // 128), ldt, device="cuda")
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
        "grad_logits_t": lambda: _lib.call("lafs_head_grad_logits_t", e_hat.data_ptr(), w_hat.data_ptr(), lab.data_ptr(), None, 1.0, B, C, D,
                                           0, 64.0, 0.4, 0, lse2.data_ptr(), go.data_ptr(), 64.0 / B, G.data_ptr(), ldg,
                                           tpart.data_ptr(), ldt, st()),
        "bwd_weight_t(dW, jac on TC)": lambda: _lib.call("lafs_head_bwd_weight_t", G.data_ptr(), ldg, e_hat.data_ptr(), w_hat.data_ptr(),
                                                         inv_w.data_ptr(), tpart.data_ptr(), tpart.shape[0], ldt, B, C, D,
                                                         dw.data_ptr(), st()),
        "bwd_embed(dE)": lambda: _lib.call("lafs_head_bwd_embed", G.data_ptr(), ldg, w_hat.data_ptr(), B, C, D, de.data_ptr(),
                                           wsb.data_ptr(), nbb, st()),
        "bwd_weight(dW+jac)": lambda: _lib.call("lafs_head_bwd_weight", G.data_ptr(), ldg, e_hat.data_ptr(), w_hat.data_ptr(),
                                                inv_w.data_ptr(), B, C, D, dw.data_ptr(), st()),
    }
    res = {"cfg": name, "B": B, "C": C, "D": D}
    for k, fn in calls.items():
        fn()
    torch.cuda.synchronize()
    sweeps = {
        "normalize_w": [{}], "normalize_e": [{}], "loss": [{}],
        "fwd_stats+merge": [{"LAFS_HEAD_1SM": "0"}, {"LAFS_HEAD_1SM": "1"}],
        "grad_logits": [{"LAFS_HEAD_1SM": "0"}, {"LAFS_HEAD_1SM": "1"}, {"LAFS_HEAD_DEBUG": "1"}],
        "bwd_embed(dE)": [{"LAFS_DE_CLUSTER": "4"}, {"LAFS_DE_CLUSTER": "2"}, {"LAFS_DE_CLUSTER": "1"}],
        "bwd_weight(dW+jac)": [{}],
        "grad_logits_t": [{}, {"LAFS_HEAD_DEBUG": "1"}],          # DEBUG=1: G stores skipped (what do the stores cost?)
        "bwd_weight_t(dW, jac on TC)": [{}],
    }
    fl = 2.0 * B * C * D
    for k, fn in calls.items():
        for env in sweeps[k]:
            os.environ.update(env)
            fn(); torch.cuda.synchronize()
            us = timeit(fn)
            tag = k + ("[" + ",".join(f"{a[5:]}={b}" for a, b in env.items()) + "]" if env else "")
            res[tag + "_us"] = us
            if "norm" not in k and k != "loss":
                res[tag + "_TFLOPs"] = round(fl / us / 1e6, 1)
            for a in env:
                os.environ.pop(a, None)
    xg = x.clone().requires_grad_(True)

    def step():
        xg.grad = None; h.weight.grad = None
        h.forward_loss(xg, lab).backward()

    def graph_us(env):
        os.environ.update(env)
        try:
            for _ in range(2):
                step()
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                step()
            for _ in range(3):
                g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            return round(e0.elapsed_time(e1) / 20 * 1e3, 1)
        finally:
            for a in env:
                os.environ.pop(a, None)

    envs = [{}, {"LAFS_HEAD_1SM": "1"}, {"LAFS_DW_DIAG": "0"}]
    for env in envs:
        tag = "step_graph[" + ",".join(f"{a[5:]}={b}" for a, b in env.items()) + "]_us"
        res[tag] = graph_us(env)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
