"""Multi-GPU (>= 2 devices, NCCL): class-sharded margin head and the DINO centre all-reduce
against the single-GPU result.  Skipped on a 1-GPU box; run with `gpurun --gpus 2`."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _guarded(rank, world, port, ret, body):
    """Run a worker body; on any failure record the traceback and leave at once (a rank that raised must not sit in
    destroy_process_group while its peer waits in a collective: that turns an assertion into a hang)."""
    import datetime
    import traceback
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank),
                            timeout=datetime.timedelta(seconds=60))
    try:
        globals()[body](rank, world, ret)
        torch.cuda.synchronize()
        ret[rank] = "ok"
    except BaseException:
        traceback.print_exc()
        ret[rank] = "FAILED: " + traceback.format_exc()[-1500:]
        os._exit(1)
    dist.destroy_process_group()


def _spawn_and_wait(body, world, port, ret, limit=200):
    import time
    ctx = mp.spawn(_guarded, args=(world, port, ret, body), nprocs=world, join=False)
    t0 = time.time()
    while time.time() - t0 < limit and any(p.is_alive() for p in ctx.processes):
        time.sleep(0.5)
    for p in ctx.processes:
        if p.is_alive():
            p.kill()


def _body_batch_sharded(rank, world, ret):
    """SURVEY 8e batch contract: each rank feeds B/world samples; all_gather(E, labels) in, reduce_scatter(dE) out."""
    import sys
    import lafs_cvpr2024_b200 as P
    torch.manual_seed(0)
    B, C, D = 192, 10007, 512
    x, w = torch.randn(B, D), torch.randn(C, D) * 0.05
    lab = torch.randint(0, C, (B,))
    lo, hi = P.shard_bounds(C, world)[rank]
    full = P.CosFace(D, C, None).cuda()
    with torch.no_grad():
        full.weight.copy_(w)
    xf = x.cuda().requires_grad_(True)
    lf = full.forward_loss(xf, lab.cuda())
    lf.backward()
    bl = B // world
    for peer in (False, True):
        print(f"[rank {rank}] batch-sharded head, peer={peer}", file=sys.stderr, flush=True)
        h3 = P.CosFace(D, C, None, shard=(rank, world), batch_sharded=True).cuda()
        if peer:
            h3.enable_peer_exchange()
        with torch.no_grad():
            h3.weight.copy_(w[lo:hi])
        x3 = x[rank * bl:(rank + 1) * bl].cuda().requires_grad_(True)
        l3 = h3.forward_loss(x3, lab[rank * bl:(rank + 1) * bl].cuda())
        print(f"[rank {rank}] forward done", file=sys.stderr, flush=True)
        l3.backward()
        torch.cuda.synchronize()
        print(f"[rank {rank}] backward done", file=sys.stderr, flush=True)
        assert abs(float(l3) - float(lf)) <= 1e-5 * abs(float(lf)), (peer, float(l3), float(lf))
        assert (x3.grad - xf.grad[rank * bl:(rank + 1) * bl]).abs().max() <= 2e-3 * xf.grad.abs().max()
        assert (h3.weight.grad - full.weight.grad[lo:hi]).abs().max() <= 2e-3 * full.weight.grad.abs().max()


def _body_main(rank, world, ret):
    if True:
        import lafs_cvpr2024_b200 as P
        torch.manual_seed(0)
        B, C, D = 192, 10007, 512
        x, w = torch.randn(B, D), torch.randn(C, D) * 0.05
        lab = torch.randint(0, C, (B,))
        lo, hi = P.shard_bounds(C, world)[rank]
        h = P.CosFace(D, C, None, shard=(rank, world)).cuda()
        with torch.no_grad():
            h.weight.copy_(w[lo:hi])
        xg = x.cuda().requires_grad_(True)
        loss = h.forward_loss(xg, lab.cuda())
        loss.backward()
        # single-GPU unsharded result on the same device
        full = P.CosFace(D, C, None).cuda()
        with torch.no_grad():
            full.weight.copy_(w)
        xf = x.cuda().requires_grad_(True)
        lf = full.forward_loss(xf, lab.cuda())
        lf.backward()
        assert abs(float(loss) - float(lf)) <= 1e-5 * abs(float(lf)), (float(loss), float(lf))
        assert (xg.grad - xf.grad).abs().max() <= 2e-3 * xf.grad.abs().max()
        assert (h.weight.grad - full.weight.grad[lo:hi]).abs().max() <= 2e-3 * full.weight.grad.abs().max()
        # the same step with the exchange through NVLink peer memory (csrc/exchange.cu) instead of NCCL
        h2 = P.CosFace(D, C, None, shard=(rank, world)).cuda().enable_peer_exchange()
        with torch.no_grad():
            h2.weight.copy_(w[lo:hi])
        for it in range(4):                          # several calls: flag epochs, slot parity
            x2 = x.cuda().requires_grad_(True)
            h2.weight.grad = None
            l2 = h2.forward_loss(x2, lab.cuda())
            l2.backward()
            assert abs(float(l2) - float(loss)) <= 1e-6 * abs(float(loss)), (it, float(l2), float(loss))
            assert (x2.grad - xg.grad).abs().max() <= 1e-5 * xg.grad.abs().max()
            assert (h2.weight.grad - h.weight.grad).abs().max() <= 1e-5 * h.weight.grad.abs().max()
        # every rank holds bit-identical summed gradients (fixed reduction order)
        gx_all = [torch.empty_like(x2.grad) for _ in range(world)]
        dist.all_gather(gx_all, x2.grad.contiguous())
        assert all(torch.equal(gx_all[0], g) for g in gx_all)
        h2._xchg[(B, D)].check()
        # DINO centre: every rank ends with the same centre = EMA of the global teacher mean
        K = 4096
        g = torch.Generator().manual_seed(10 + rank)
        t = torch.randn(8, K, generator=g).cuda()
        s = torch.randn(16, K, generator=g).cuda()
        crit = P.DINOLoss(K, 4, 0.04, 0.07, 30, 41).cuda()
        crit(s, t, 0)
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        ref = torch.cat(allt).sum(0, keepdim=True) / (8 * world) * (1 - 0.9)
        torch.testing.assert_close(crit.center, ref, rtol=1e-5, atol=1e-6)
        # the same centre update with the column sums all-reduced through peer memory; single-call form too
        crit2 = P.DINOLoss(K, 4, 0.04, 0.07, 30, 41).cuda().enable_peer_exchange()
        crit2(s, t, 0)
        torch.testing.assert_close(crit2.center, crit.center, rtol=1e-6, atol=1e-7)
        crit3 = P.DINOLoss(K, 4, 0.04, 0.07, 30, 41).cuda().enable_peer_exchange()
        for _ in range(3):
            crit3.center = torch.zeros(1, K, device="cuda")
            l3, g3 = crit3.loss_and_grad(s.bfloat16(), t.bfloat16(), 0)
        cref = P.DINOLoss(K, 4, 0.04, 0.07, 30, 41).cuda()
        cref.loss_and_grad(s.bfloat16(), t.bfloat16(), 0)
        torch.testing.assert_close(crit3.center, cref.center, rtol=1e-6, atol=1e-7)
        cs = [torch.empty_like(crit3.center) for _ in range(world)]
        dist.all_gather(cs, crit3.center.contiguous())
        assert all(torch.equal(cs[0], c) for c in cs)      # identical bits on every rank


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sharded_head_and_center_two_gpus():
    world = 2
    port = 29800 + (os.getpid() % 100)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        _spawn_and_wait("_body_main", world, port, ret)
        assert dict(ret) == {0: "ok", 1: "ok"}, dict(ret)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_batch_sharded_head_two_gpus():
    world = 2
    port = 29900 + (os.getpid() % 100)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        _spawn_and_wait("_body_batch_sharded", world, port, ret)
        assert dict(ret) == {0: "ok", 1: "ok"}, dict(ret)
