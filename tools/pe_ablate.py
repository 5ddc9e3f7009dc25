import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = '''
import sys, torch
sys.path.insert(0, %r)
import lafs_cvpr2024_b200 as P
torch.manual_seed(0)
th = torch.rand(512, 196, 2, device="cuda") * 111
thl = torch.rand(1024, 36, 2, device="cuda") * 111
a, b = torch.nn.Linear(192, 768).cuda(), torch.nn.Linear(192, 768).cuda()
w2 = P.PatchEmbedWeights([(a.weight, a.bias), (b.weight, b.bias)]); w1 = P.PatchEmbedWeights([(a.weight, a.bias)])
u8g = torch.randint(0, 256, (512, 3, 112, 112), dtype=torch.uint8, device="cuda")
u8l = torch.randint(0, 256, (1024, 3, 112, 112), dtype=torch.uint8, device="cuda")
fg = torch.rand(512, 3, 112, 112, device="cuda")
def t(fn):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 20 * 1000
print("u8_global %%.1f us  u8_local %%.1f us  f32_global %%.1f us" %% (t(lambda: P.gather_embed(u8g, th, w2)), t(lambda: P.gather_embed(u8l, thl, w1)), t(lambda: P.gather_embed(fg, th, w2))))
''' % ROOT
for dbg in (0, 1, 2, 4, 3, 5, 6, 7):
    env = dict(os.environ, LAFS_PE_DEBUG=str(dbg))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print("debug=%d (1=no stores 2=no gather 4=no mma):" % dbg, r.stdout.strip(), r.stderr.strip()[-200:])
