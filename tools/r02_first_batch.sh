#!/bin/bash
# First GPU call of round 2: verify and measure the two experimental variants written after round 1's GPU
# budget was spent (each under a hard timeout, in its own process: a hung kernel must not eat the budget).
#   LAFS_PE_DEEP_RING=1  7-slot plane ring + one token tile for uint8 views with <= 64 landmarks (patch_embed.cu)
#   LAFS_DW_DIAG=1       dW Jacobian on the tensor core (head.cu HEAD_GRAD_T + head_bwd.cu dw_diag_kernel)
mkdir -p gpurun_out
T="timeout -s KILL"
export PYTHONUNBUFFERED=1 LAFS_TEST_EXPERIMENTAL=1
echo "=== deep plane ring: parity"; $T 120 python -m pytest tests/test_gpu_patch_embed.py -q -k deep_plane_ring -p no:cacheprovider 2>&1 | tail -3
echo "=== dW diag: parity"; $T 240 python -m pytest tests/test_gpu_head.py -q -k tensor_core_variant -p no:cacheprovider 2>&1 | tail -5
echo "=== patch-embed local views, default vs deep ring"
for v in 0 1; do
  LAFS_PE_DEEP_RING=$v $T 120 python - <<'PY'
import os, torch, sys
sys.path.insert(0, os.getcwd())
import lafs_cvpr2024_b200 as P
torch.manual_seed(0)
u8 = torch.randint(0, 256, (1024, 3, 112, 112), dtype=torch.uint8, device="cuda")
th = torch.rand(1024, 36, 2, device="cuda") * 111
lin = torch.nn.Linear(192, 768).cuda()
w = P.PatchEmbedWeights([(lin.weight, lin.bias)])
for _ in range(5):
    P.gather_embed(u8, th, w)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
for a, b in ev:
    a.record(); P.gather_embed(u8, th, w); b.record()
torch.cuda.synchronize()
ts = sorted(a.elapsed_time(b) for a, b in ev)
print("LAFS_PE_DEEP_RING=%s  gather_embed u8 1024x36: %.1f us (median of 20)" % (os.environ.get("LAFS_PE_DEEP_RING"), ts[10] * 1e3))
PY
done
echo "=== head step, default vs LAFS_DW_DIAG=1 (graph timings at the end of each line)"
for c in cfg3 cfg4; do $T 300 python tools/head_breakdown.py $c | tail -1 | tee -a gpurun_out/head_breakdown_r02.jsonl; done
echo "=== write-bandwidth probe (SIMT vs TMA bulk store vs memset)"
$T 120 tools/_bin/wbw_probe 2>&1 | tee gpurun_out/wbw_probe.txt
echo "=== full gpu suite (experimental tests on)"
$T 900 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -5
