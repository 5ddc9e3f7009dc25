"""CPU, world_size 2, gloo: the host-side logic of the N>1 path -- class-shard ownership
(torch.chunk rule), the per-row softmax-statistics exchange of the class-parallel head, and the
all-reduced DINO centre update -- checked with real collectives."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import lafs_oracle as O


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from lafs_cvpr2024_b200.margin_head import label_to_shard, shard_bounds
        torch.manual_seed(0)                       # same global data on every rank
        B, C, D = 16, 1003, 32
        x, w = torch.randn(B, D), torch.randn(C, D)
        lab = torch.randint(0, C, (B,))
        lab[0], lab[1] = 0, C - 1
        # (a) ownership: ranks' ranges partition [0, C) exactly like torch.chunk
        lo, hi = shard_bounds(C, world)[rank]
        sizes = [None] * world
        dist.all_gather_object(sizes, (lo, hi))
        assert sizes[0][0] == 0 and sizes[-1][1] == C
        assert all(sizes[i][1] == sizes[i + 1][0] for i in range(world - 1))
        assert [h - l for l, h in sizes] == [c.shape[0] for c in torch.chunk(w, world, dim=0)]
        owner, local = label_to_shard(lab, C, world)
        mine = owner == rank
        assert torch.equal(lab[mine] - lo, local[mine])
        # (b) statistics exchange: all_gather of per-shard (max, sumexp, target logit) == full softmax CE
        logits_full = O.cosface_logits(x, w, lab)
        z_local = logits_full[:, lo:hi]
        m, l = O.softmax_stats(z_local)
        tgt = torch.where(mine, logits_full[torch.arange(B), lab], torch.zeros(B))
        rec = torch.stack([m, l, tgt], 1)
        parts = [torch.empty_like(rec) for _ in range(world)]
        dist.all_gather(parts, rec)
        lse = O.merge_softmax_stats([p[:, 0] for p in parts], [p[:, 1] for p in parts])
        loss = (lse - sum(p[:, 2] for p in parts)).mean()
        ref = O.cross_entropy(logits_full, lab)
        assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
        # (c) DINO centre: all-reduced column sums / (rows * world)  (lafs_train.py:674-679)
        g = torch.Generator().manual_seed(100 + rank)
        t_local = torch.randn(6, 64, generator=g)
        colsum = t_local.sum(0, keepdim=True)
        dist.all_reduce(colsum)
        center = O.dino_center_update(torch.zeros(1, 64), t_local, world_size=world, allreduced_sum=colsum)
        all_t = [torch.empty_like(t_local) for _ in range(world)]
        dist.all_gather(all_t, t_local)
        ref_c = torch.cat(all_t).sum(0, keepdim=True) / (6 * world) * (1 - 0.9)
        torch.testing.assert_close(center, ref_c, rtol=1e-6, atol=1e-7)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_world2_gloo_host_logic():
    world = 2
    port = 29600 + (os.getpid() % 200)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: "ok", 1: "ok"}


def test_bench_reference_arm_under_two_ranks():
    """bench.py --impl reference: rank 0 prints the line, the other rank exits 0 without work."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
