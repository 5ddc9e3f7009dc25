#!/bin/bash
# GPU batch 2 of round 2: full suite on the new kernels (DINO pre-merged records + PDL finish, token-major patch-embed,
# dW Jacobian on the tensor core by default), A/B timings, bench (ours with reference_eager_b200, reference arm).
mkdir -p gpurun_out
T="timeout -s KILL"
export PYTHONUNBUFFERED=1
echo "=== pytest -m gpu"; $T 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -8
echo "=== patch-embed A/B (token-major default vs LAFS_PE_TOKN=1)"
for v in 0 1; do
  LAFS_PE_TOKN=$v $T 200 python - <<'PY'
import os, torch, sys
sys.path.insert(0, os.getcwd())
import lafs_cvpr2024_b200 as P
torch.manual_seed(0)
la, lb = torch.nn.Linear(192, 768).cuda(), torch.nn.Linear(192, 768).cuda()
w2 = P.PatchEmbedWeights([(la.weight, la.bias), (lb.weight, lb.bias)])
w1 = P.PatchEmbedWeights([(la.weight, la.bias)])
def t(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[n // 2] * 1e3
for name, dt in (("u8", torch.uint8), ("f32", torch.float32)):
    g = torch.randint(0, 256, (512, 3, 112, 112), dtype=torch.uint8, device="cuda")
    l = torch.randint(0, 256, (1024, 3, 112, 112), dtype=torch.uint8, device="cuda")
    if dt == torch.float32:
        g, l = g.float() / 127.5 - 1, l.float() / 127.5 - 1
    thg = torch.rand(512, 196, 2, device="cuda") * 111
    thl = torch.rand(1024, 36, 2, device="cuda") * 111
    print("TOKN=%s %s  global 512x196 x2 models: %.1f us   local 1024x36: %.1f us" % (
        os.environ.get("LAFS_PE_TOKN"), name, t(lambda: P.gather_embed(g, thg, w2)), t(lambda: P.gather_embed(l, thl, w1))))
PY
done
echo "=== DINO forward: PDL on/off"
for v in 1 0; do
  LAFS_DINO_PDL=$v $T 200 python - <<'PY'
import os, torch, sys
sys.path.insert(0, os.getcwd())
import lafs_cvpr2024_b200 as P
torch.manual_seed(0)
B, K, nc = 256, 65536, 6
s = torch.randn(nc * B, K, device="cuda").bfloat16(); t = torch.randn(2 * B, K, device="cuda").bfloat16()
crit = P.DINOLoss(K, nc, 0.04, 0.07, 30, 41).cuda()
def run(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[n // 2] * 1e3
with torch.no_grad():
    f = run(lambda: crit(s, t, 3))
fb = run(lambda: crit.loss_and_grad(s, t, 3))
print("PDL=%s  dino fwd+centre %.1f us (%.3f of 6453.7 GB/s)   fwd+bwd %.1f us (%.3f)" % (
    os.environ.get("LAFS_DINO_PDL"), f, 269e6 / f / 6453.7e3, fb, 739e6 / fb / 6453.7e3))
PY
done
echo "=== head breakdown"
for c in cfg3 cfg4; do $T 300 python tools/head_breakdown.py $c | tail -1 | tee -a gpurun_out/head_breakdown_r02b.jsonl | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:v for k,v in d.items() if k.endswith('_us') or k=='cfg'})"; done
echo "=== bench ours"; $T 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.err; head -c 1500 gpurun_out/bench_n1.json; echo
echo "=== bench reference"; $T 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; head -c 900 gpurun_out/bench_ref.json; echo
