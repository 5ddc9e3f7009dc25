#!/bin/bash
# second focused pass for the fused DINO head: side-stream overlap on/off, dE cluster width, its tests, the new head / ViT tests
mkdir -p gpurun_out
T="timeout -s KILL"
export PYTHONUNBUFFERED=1
echo "=== pytest (dino head, new head / vit tests)"
$T 600 python -m pytest tests/test_gpu_dino_head.py tests/test_gpu_vit_face.py "tests/test_gpu_head.py::test_arcface_forward_logits_backward_vs_oracle" \
   "tests/test_gpu_head.py::test_label_range_check_is_opt_in" "tests/test_gpu_head.py::test_head_mixup_soft_labels_vs_oracle" \
   "tests/test_gpu_head.py::test_head_backward_soft_labels_and_grad_out" -q -p no:cacheprovider --tb=short 2>&1 | tail -40
export LAFS_PROBE_FUSED_ONLY=1
for ov in 1 0; do for cl in 4 2 1; do
  echo "--- overlap=$ov de_cluster=$cl"
  LAFS_DH_OVERLAP=$ov LAFS_DE_CLUSTER=$cl $T 120 python tools/dino_head_probe.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['fused']['ms_fwd'], d['fused']['ms_fwd_bwd'])"
done; done | tee gpurun_out/f1_sweep.txt
