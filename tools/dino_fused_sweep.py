import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = '''
import sys, torch
sys.path.insert(0, %r)
import lafs_cvpr2024_b200 as P
B, K, nc = 256, 65536, 6
torch.manual_seed(0)
s = torch.randn(nc * B, K, device="cuda", dtype=torch.bfloat16); t = torch.randn(2 * B, K, device="cuda", dtype=torch.bfloat16)
crit = P.DINOLoss(K, nc, 0.04, 0.07, 30, 41).cuda()
def t_(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1000
print("fused fwd+bwd %%.1f us" %% t_(lambda: crit.loss_and_grad(s, t, 3)))
''' % ROOT
for wave, groups in ((0, 0), (256, 0), (128, 0), (64, 0), (64, 9), (48, 0), (32, 0), (32, 9)):
    env = dict(os.environ)
    if wave: env["LAFS_DINO_WAVE"] = str(wave)
    if groups: env["LAFS_DINO_GROUPS"] = str(groups)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("wave=%d groups=%d:" % (wave, groups), r.stdout.strip(), r.stderr.strip()[-300:])
