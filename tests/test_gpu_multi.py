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


def _body_fused_dino_head(rank, world, ret):
    """(f1) fused last_layer + DINO loss on several ranks: the loss is rank-local (DDP averages gradients), the centre
    moves by the EMA of the GLOBAL teacher mean (all-reduce of the [K] column sums, lafs_train.py:675) and ends with
    identical bits on every rank -- through NCCL and through the peer-memory all-reduce."""
    import lafs_cvpr2024_b200 as P
    in_dim, K, B, nc = 64, 2048, 8, 4
    torch.manual_seed(0)                                     # same heads on every rank
    hs = P.DINOHead(in_dim, K, nlayers=2, hidden_dim=96, bottleneck_dim=64, fused_loss=True).cuda()
    ht = P.DINOHead(in_dim, K, nlayers=2, hidden_dim=96, bottleneck_dim=64, fused_loss=True).cuda()
    with torch.no_grad():
        for h in (hs, ht):
            h.last_layer.weight_v.normal_(0, 0.5)
    g = torch.Generator().manual_seed(20 + rank)             # different samples per rank
    fs, ft = torch.randn(nc * B, in_dim, generator=g).cuda(), torch.randn(2 * B, in_dim, generator=g).cuda()
    c0 = torch.randn(1, K, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3)) * 0.05
    centres = []
    for peer in (False, True):
        crit = P.DINOLoss(K, nc, 0.04, 0.07, 30, 41).cuda()
        if peer:
            crit.enable_peer_exchange()
        crit.center = c0.clone()
        with torch.no_grad():
            to = ht(ft)
        loss = crit(hs(fs), to, 3)
        loss.backward()
        # reference: the unfused modules on this rank's samples + the global teacher mean
        with torch.no_grad():
            t_log = to.logits()
        allt = [torch.empty_like(t_log) for _ in range(world)]
        dist.all_gather(allt, t_log.contiguous())
        ref_c = c0 * 0.9 + torch.cat(allt).float().sum(0, keepdim=True) / (2 * B * world) * (1 - 0.9)
        torch.testing.assert_close(crit.center, ref_c, rtol=0, atol=2e-3 * float(ref_c.abs().max()) + 1e-4)
        cu = P.DINOLoss(K, nc, 0.04, 0.07, 30, 41).cuda()
        cu.center = c0.clone()
        hs.fused_loss = False
        lu = cu(hs(fs), t_log, 3)
        hs.fused_loss = True
        assert abs(float(loss) - float(lu)) <= 1e-3 * abs(float(lu)), (peer, float(loss), float(lu))
        cs = [torch.empty_like(crit.center) for _ in range(world)]
        dist.all_gather(cs, crit.center.contiguous())
        assert all(torch.equal(cs[0], c) for c in cs)        # identical bits on every rank
        centres.append(crit.center.clone())
    torch.testing.assert_close(centres[0], centres[1], rtol=1e-6, atol=1e-7)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_fused_dino_head_two_gpus():
    world = 2
    port = 29700 + (os.getpid() % 100)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        _spawn_and_wait("_body_fused_dino_head", world, port, ret)
        assert dict(ret) == {0: "ok", 1: "ok"}, dict(ret)


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
