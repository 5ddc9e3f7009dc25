import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = '''
import sys, torch
sys.path.insert(0, %r)
from lafs_cvpr2024_b200 import _lib
B, K, nc = 256, 65536, 6
s = torch.randn(nc * B, K, device="cuda", dtype=torch.bfloat16); t = torch.randn(2 * B, K, device="cuda", dtype=torch.bfloat16)
c = torch.randn(K, device="cuda") * 0.1
loss = torch.empty((), device="cuda"); rs = torch.empty((nc + 2) * B, device="cuda"); cs = torch.empty(K, device="cuda")
nb = _lib.lib().lafs_dino_workspace_bytes(B, K, nc); ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
f = lambda: _lib.call("lafs_dino_fwd", s.data_ptr(), t.data_ptr(), c.data_ptr(), B, K, nc, 10.0, 25.0, 1, loss.data_ptr(), rs.data_ptr(), cs.data_ptr(), ws.data_ptr(), nb, None, 0.0, 0.0, _lib.stream())
for _ in range(5): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(30): f()
e1.record(); torch.cuda.synchronize()
print("%%.1f us  loss %%.6f" %% (e0.elapsed_time(e1) / 30 * 1000, float(loss)))
''' % ROOT
for g in (0,):  # groups
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, LAFS_DINO_GROUPS=str(g)), capture_output=True, text=True)
    print("groups=%d:" % g, r.stdout.strip(), r.stderr.strip()[-200:])
