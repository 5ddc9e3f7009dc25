"""One eager step of the fused DINO head or of its unfused form, for ncu launch lists.
    python tools/dino_head_once.py {fused|unfused}      (BASELINE configs[1] size)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lafs_cvpr2024_b200 as P  # noqa: E402

mode = sys.argv[1]
B, ncrops, K, D = 256, 6, 65536, 256
dev = torch.device("cuda", 0)
torch.manual_seed(11)
xs = torch.randn(ncrops * B, D, device=dev)
xt = torch.randn(2 * B, D, device=dev)
vs = torch.randn(K, D, device=dev) * 0.02
vt = vs + torch.randn(K, D, device=dev) * 0.002
one = torch.ones(K, device=dev)
center = torch.randn(K, device=dev) * 0.05
torch.cuda.synchronize()
torch.cuda.profiler.start()          # ncu --profile-from-start off: the input generation above is not part of the step
if mode == "fused":
    loss, colsum, saved = P.dino_head_forward(xs, xt, vs, one, vt, one, center, ncrops, 10.0, 25.0)
    P.dino_head_backward(saved, torch.ones((), device=dev))
else:
    dl = P.DINOLoss(K, ncrops, 0.04, 0.04, 0, 1).to(dev)
    dl.center = center.view(1, -1).clone()
    xs_u = xs.clone().requires_grad_(True)
    vs_u = vs.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ws = vs_u * (one / vs_u.norm(dim=1)).unsqueeze(1)
        s_out = torch.nn.functional.linear(torch.nn.functional.normalize(xs_u, dim=-1, p=2), ws)
        with torch.no_grad():
            wt = vt * (one / vt.norm(dim=1)).unsqueeze(1)
            t_out = torch.nn.functional.linear(torch.nn.functional.normalize(xt, dim=-1, p=2), wt)
    loss = dl(s_out, t_out, 0)
    loss.backward()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(mode, float(loss))
