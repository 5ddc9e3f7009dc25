"""Launches one hot-path kernel a few times so ncu can capture it (development aid).
usage: python tools/prof_driver.py {pe_global|pe_local|dino|head|ema|gather}"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lafs_cvpr2024_b200 as P  # noqa: E402
from lafs_cvpr2024_b200 import _lib  # noqa: E402

mode = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
if mode in ("pe_global", "pe_local", "pe_global_u8", "pe_local_u8"):
    g = not mode.startswith("pe_local")
    Bv, n = (512, 196) if g else (1024, 36)
    imgs = torch.rand(Bv, 3, 112, 112, device="cuda") * 2 - 1
    if mode.endswith("_u8"):
        imgs = torch.randint(0, 256, (Bv, 3, 112, 112), dtype=torch.uint8, device="cuda")
    th = torch.rand(Bv, n, 2, device="cuda") * 111
    a, b = torch.nn.Linear(192, 768).cuda(), torch.nn.Linear(192, 768).cuda()
    w = P.PatchEmbedWeights([(a.weight, a.bias), (b.weight, b.bias)] if g else [(a.weight, a.bias)])
    for _ in range(reps):
        P.gather_embed(imgs, th, w)
elif mode == "dino":
    B, K, nc = 256, 65536, 6
    s = torch.randn(nc * B, K, device="cuda", dtype=torch.bfloat16).requires_grad_(True)
    t = torch.randn(2 * B, K, device="cuda", dtype=torch.bfloat16)
    crit = P.DINOLoss(K, nc, 0.04, 0.07, 30, 41).cuda()
    for _ in range(reps):
        s.grad = None
        crit(s, t, 3).backward()
elif mode == "head":
    B, C, D = 512, 93431, 512
    h = P.CosFace(D, C, None).cuda()
    x = torch.randn(B, D, device="cuda")
    lab = torch.randint(0, C, (B,), device="cuda")
    for _ in range(reps):
        h.forward_loss(x, lab)
elif mode == "head_bwd":
    B, C, D = 512, 93431, 512
    h = P.CosFace(D, C, None).cuda()
    x = torch.randn(B, D, device="cuda").requires_grad_(True)
    lab = torch.randint(0, C, (B,), device="cuda")
    for _ in range(reps):
        x.grad = None; h.weight.grad = None
        h.forward_loss(x, lab).backward()
elif mode == "ema":
    q = [torch.randn(30000, 768, device="cuda"), torch.randn(65536, 256, device="cuda")] + [torch.randn(2112, 768, device="cuda") for _ in range(12)]
    k = [a.clone() for a in q]
    for _ in range(reps):
        P.ema_update_(k, q, 0.996)
elif mode == "optim":
    import bench
    shapes = bench.vit_param_shapes("B")
    p = [torch.randn(*s, device="cuda") * 0.02 for s in shapes]
    k = [a.clone() for a in p]
    gr = [torch.randn_like(a) * 0.01 for a in p]
    upd = P.StudentUpdate(p, k, regularized=[a.dim() > 1 for a in p])
    for _ in range(reps):
        upd.step(gr, 5e-4, 0.04, 3.0, 0.996)
elif mode == "pe_bwd":
    M, dim = 100352, 768
    dy = (torch.randn(M, dim, device="cuda") * 0.01).bfloat16()
    tok = P.new_token_buffer(M, "cuda")
    tok[:, :192] = torch.randn(M, 192, device="cuda").bfloat16()
    for _ in range(reps):
        P.embed_backward_weight(dy, tok)
elif mode == "gather":
    imgs = torch.rand(512, 3, 112, 112, device="cuda") * 2 - 1
    th = torch.rand(512, 196, 2, device="cuda") * 111
    for _ in range(reps):
        P.extract_tokens(imgs, th)
torch.cuda.synchronize()
print("done", mode)
