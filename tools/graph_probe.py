"""Does the whole SSL hot-path step capture into a CUDA graph, and what does it buy?"""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from lafs_cvpr2024_b200.ssl_step import SSLHotPath, GraphedSSLStep
dev = torch.device("cuda", 0)
B, L = bench.B_PER_GPU, bench.N_LOCAL
host = bench.make_host_inputs(B, 1, uint8=True)
st = bench.make_device_state(B, 2, dev)
inp = {k: v.to(dev) for k, v in host.items()}
inp.update(raw_g=st["raw_g"], raw_l=st["raw_l"], student_out=st["student_out"], teacher_out=st["teacher_out"],
           grad_s_g=st["grad_s_g"], grad_s_l=st["grad_s_l"])
path = SSLHotPath(bench.OUT_DIM, L, st["teacher_params"], st["student_params"],
                  student_embed=(st["student_params"][1], st["student_params"][2]),
                  teacher_embed=(st["teacher_params"][1], st["teacher_params"][2]))
path.loss.center = torch.randn(1, bench.OUT_DIM, device=dev) * 0.1
import sys as _s
mode = os.environ.get("EMA_OVERLAP", "")          # "", "early", "late"
g = GraphedSSLStep(path, inp, epoch=3, momentum=0.996, overlap_ema=(mode or False),
                   ema_ctas=int(os.environ.get("EMA_CTAS", "148")))
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
c0 = g.center.clone()
l1 = float(g.replay()); c1 = g.center.clone()
l2 = float(g.replay()); c2 = g.center.clone()
print("loss", l1, l2, "centre moved:", float((c1 - c0).abs().max()), float((c2 - c1).abs().max()))
print("EMA_OVERLAP=%r EMA_CTAS=%s graph replay ms/step: %.4f" % (mode, os.environ.get("EMA_CTAS", "148"), t(g.replay)))
