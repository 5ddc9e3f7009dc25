import sys, torch
sys.path.insert(0, ".")
import lafs_cvpr2024_b200 as P
torch.manual_seed(0)
thl = torch.rand(1024, 36, 2, device="cuda") * 111
a = torch.nn.Linear(192, 768).cuda()
w1 = P.PatchEmbedWeights([(a.weight, a.bias)])
u8l = torch.randint(0, 256, (1024, 3, 112, 112), dtype=torch.uint8, device="cuda")
def t(fn):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 30 * 1000
print("u8_local %.1f us" % t(lambda: P.gather_embed(u8l, thl, w1)))
