"""Times the fused DINO head against the unfused form on the same GPU; one JSON line.
    python tools/dino_head_probe.py [B ncrops K D]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:]]
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    fused_only = os.environ.get("LAFS_PROBE_FUSED_ONLY", "0") not in ("", "0")
    print(json.dumps(bench.bench_dino_head(dev, 20, *a, fused_only=fused_only)))
