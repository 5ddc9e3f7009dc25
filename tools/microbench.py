"""Per-kernel timings on one B200 (development aid; bench.py is the judged entry point).
Prints achieved GB/s against MEASURED_PEAKS.json for the HBM-bound kernels."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lafs_cvpr2024_b200 as P  # noqa: E402
from lafs_cvpr2024_b200 import _lib  # noqa: E402


def timeit(fn, warmup=5, iters=20):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2], ts[0]


def main():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    res = {}
    torch.manual_seed(0)
    # ---- EMA: ViT-B student+head parameter list (147 tensors, ~110M params)
    shapes = [(1, 197, 768), (1, 1, 768), (768, 192), (768,), (30000, 768)]
    for _ in range(12):
        shapes += [(768,), (768,), (2112, 768), (768, 704), (768,), (768,), (768,), (2048, 768), (2048,), (768, 2048), (768,)]
    shapes += [(768,), (768,), (2048, 768), (2048,), (2048, 2048), (2048,), (256, 2048), (256,), (65536, 1), (65536, 256)]
    q = [torch.randn(*s, device="cuda") for s in shapes]
    k = [torch.randn(*s, device="cuda") for s in shapes]
    nparam = sum(a.numel() for a in q)
    plan = P.ema_update_(k, q, 0.996)
    med, best = timeit(lambda: plan.step(0.996))
    res["ema"] = {"tensors": len(shapes), "params": nparam, "ms": med, "ms_best": best,
                  "GBps": 12 * nparam / med / 1e6, "frac_hbm": 12 * nparam / med / 1e6 / hbm}
    del q, k
    # ---- DINO config 2: B=256, K=65536, ncrops=6, bf16
    for (B, K, ncrops) in [(256, 65536, 6), (256, 65536, 10)]:
        s = torch.randn(ncrops * B, K, device="cuda", dtype=torch.bfloat16)
        t = torch.randn(2 * B, K, device="cuda", dtype=torch.bfloat16)
        c = torch.randn(K, device="cuda") * 0.1
        loss = torch.empty((), device="cuda"); rs = torch.empty((ncrops + 2) * B, device="cuda"); cs = torch.empty(K, device="cuda")
        nb = _lib.lib().lafs_dino_workspace_bytes(B, K, ncrops)
        ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
        gs = torch.empty_like(s); go = torch.ones((), device="cuda")
        f = lambda: _lib.call("lafs_dino_fwd", s.data_ptr(), t.data_ptr(), c.data_ptr(), B, K, ncrops, 10.0, 25.0, 1,
                              loss.data_ptr(), rs.data_ptr(), cs.data_ptr(), ws.data_ptr(), nb, None, 0.0, 0.0, _lib.stream())
        b = lambda: _lib.call("lafs_dino_bwd", s.data_ptr(), t.data_ptr(), c.data_ptr(), rs.data_ptr(), go.data_ptr(),
                              B, K, ncrops, 10.0, 25.0, 1, gs.data_ptr(), _lib.stream())
        mf, bf = timeit(f)
        mb, bb = timeit(b)
        fb = (ncrops + 2) * B * K * 2 + 8 * K
        bbytes = (2 * ncrops + 2) * B * K * 2
        res[f"dino_B{B}_K{K}_c{ncrops}"] = {
            "fwd_ms": mf, "fwd_GBps": fb / mf / 1e6, "fwd_frac": fb / mf / 1e6 / hbm,
            "bwd_ms": mb, "bwd_GBps": bbytes / mb / 1e6, "bwd_frac": bbytes / mb / 1e6 / hbm}
        del s, t, gs
    # ---- stand-alone gather, config 2 globals: 512 faces x 196 landmarks -> tokens fp32
    imgs = torch.rand(512, 3, 112, 112, device="cuda") * 2 - 1
    th = torch.rand(512, 196, 2, device="cuda") * 111
    out = torch.empty(512, 196, 192, device="cuda")
    g = lambda: _lib.call("lafs_gather_fwd", imgs.data_ptr(), th.data_ptr(), out.data_ptr(), 512, 3, 112, 112, 196, 1, 0, _lib.stream())
    mg, bg = timeit(g)
    gb = imgs.numel() * 4 + out.numel() * 4
    res["gather_512x196"] = {"ms": mg, "GBps": gb / mg / 1e6, "frac_hbm": gb / mg / 1e6 / hbm}
    # ---- fused gather -> patch embed (config 2: 512 global faces x2 models, 1024 local faces)
    tcp = peaks.get("bf16_tflops_sustained", 1400.0)
    lin_s, lin_t = torch.nn.Linear(192, 768).cuda(), torch.nn.Linear(192, 768).cuda()
    w2 = P.PatchEmbedWeights([(lin_s.weight, lin_s.bias), (lin_t.weight, lin_t.bias)])
    w1 = P.PatchEmbedWeights([(lin_s.weight, lin_s.bias)])
    imgs_l = torch.rand(1024, 3, 112, 112, device="cuda") * 2 - 1
    th_l = torch.rand(1024, 36, 2, device="cuda") * 111
    mg2, _ = timeit(lambda: P.gather_embed(imgs, th, w2))
    ml1, _ = timeit(lambda: P.gather_embed(imgs_l, th_l, w1))
    u8g = torch.randint(0, 256, (512, 3, 112, 112), dtype=torch.uint8, device="cuda")
    u8l = torch.randint(0, 256, (1024, 3, 112, 112), dtype=torch.uint8, device="cuda")
    mg2u, _ = timeit(lambda: P.gather_embed(u8g, th, w2))
    ml1u, _ = timeit(lambda: P.gather_embed(u8l, th_l, w1))
    res["gather_embed_u8_global"] = {"ms": mg2u, "GBps": (by_g0 := 512 * (3 * 112 * 112 + 196 * 8) + 2 * 512 * 196 * 768 * 2) / mg2u / 1e6}
    res["gather_embed_u8_local"] = {"ms": ml1u, "GBps": (1024 * (3 * 112 * 112 + 36 * 8) + 1024 * 36 * 768 * 2) / ml1u / 1e6}
    del u8g, u8l
    by_g = 512 * (3 * 112 * 112 * 4 + 196 * 8) + 2 * 512 * 196 * 768 * 2
    fl_g = 2.0 * 192 * 768 * 2 * 512 * 196
    by_l = 1024 * (3 * 112 * 112 * 4 + 36 * 8) + 1024 * 36 * 768 * 2
    fl_l = 2.0 * 192 * 768 * 1024 * 36
    res["gather_embed_global_512x196x2models"] = {"ms": mg2, "GBps": by_g / mg2 / 1e6, "frac_hbm": by_g / mg2 / 1e6 / hbm,
                                                  "TFLOPs": fl_g / mg2 / 1e9, "frac_tc": fl_g / mg2 / 1e9 / tcp}
    res["gather_embed_local_1024x36"] = {"ms": ml1, "GBps": by_l / ml1 / 1e6, "frac_hbm": by_l / ml1 / 1e6 / hbm,
                                         "TFLOPs": fl_l / ml1 / 1e9, "frac_tc": fl_l / ml1 / 1e9 / tcp}
    del imgs_l
    if "--head" in sys.argv or True:
        tc = peaks.get("bf16_tflops_sustained", 1400.0)
        for (B, C, D) in [(512, 93431, 512), (1024, 205990, 512), (512, 93431, 768)]:
            h = P.CosFace(D, C, None).cuda()
            x = torch.randn(B, D, device="cuda")
            lab = torch.randint(0, C, (B,), device="cuda")
            st, (la, lb, lam, e_hat, w_hat) = h.forward_stats(x, lab)
            nb = _lib.lib().lafs_head_workspace_bytes(B, C, D)
            ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
            f = lambda: _lib.call("lafs_head_fwd", e_hat.data_ptr(), w_hat.data_ptr(), la.data_ptr(), None, 1.0, B, C, D, 0,
                                  64.0, 0.4, 0, st.data_ptr(), ws.data_ptr(), nb, _lib.stream())
            mf, bf = timeit(f)
            prep = lambda: _lib.call("lafs_normalize_rows", h.weight.data_ptr(), 0, C, D, w_hat.data_ptr(), None, _lib.stream())
            mp, bp = timeit(prep)
            fl = 2.0 * B * C * D
            xg = x.clone().requires_grad_(True)
            def fb():
                xg.grad = None; h.weight.grad = None
                h.forward_loss(xg, lab).backward()
            mfb, _ = timeit(fb, warmup=3, iters=10)
            res[f"head_fwd_B{B}_C{C}_D{D}"] = {"ms": mf, "TFLOPs": fl / mf / 1e9, "frac_tc": fl / mf / 1e9 / tc,
                                               "w_prep_ms": mp, "w_prep_GBps": C * D * 6 / mp / 1e6,
                                               "fwd_bwd_step_ms": mfb, "fwd_bwd_TFLOPs_6BCD": 3 * fl / mfb / 1e9,
                                               "fwd_bwd_faces_per_s": B / mfb * 1e3}
            del h, w_hat
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
