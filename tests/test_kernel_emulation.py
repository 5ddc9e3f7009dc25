"""CPU: the CUDA gather kernel's op sequence (restated in numpy) is bit-identical to the
oracle / reference, in both coordinate modes."""
import torch

from oracle import lafs_oracle as O
from tests.kernel_emulation import gather_tokens


def test_gather_sequence_bit_exact():
    torch.manual_seed(0)
    B = 4
    imgs = torch.rand(B, 3, 112, 112) * 2 - 1
    th = torch.rand(B, 196, 2) * 111 + torch.randn(B, 196, 2) * 5
    for recip in (False, True):
        ref = O.extract_tokens(imgs, th, recip_mul=recip).reshape(B, 196, 8, 8, 3).numpy()
        emu = gather_tokens(imgs.numpy(), th.numpy(), recip=recip)
        assert (emu == ref).all(), recip
    # the two coordinate modes genuinely differ by more than the 1e-5 tolerance
    a = O.extract_tokens(imgs, th, recip_mul=False)
    b = O.extract_tokens(imgs, th, recip_mul=True)
    assert (a - b).abs().max() > 1e-5


def test_peer_allreduce_slices_partition_the_array():
    """Every float4 of the array is summed by exactly one thread of exactly one rank, for ragged sizes too."""
    from tests.kernel_emulation import xchg_slices
    for n4, world in ((65536, 8), (262144 // 16, 2), (1000, 8), (7, 8), (16384, 4), (3, 2)):
        sl = xchg_slices(n4, world, nctas=4, threads=32)
        flat = [i for s_ in sl for i in s_]
        assert sorted(flat) == list(range(n4)), (n4, world)


def test_peer_allreduce_emulation_equals_sum():
    import numpy as np
    from tests.kernel_emulation import xchg_allreduce
    rng = np.random.default_rng(0)
    for world, shape in ((2, (16, 64)), (8, (5, 12)), (4, (1, 4096))):
        parts = [rng.standard_normal(shape).astype(np.float32) for _ in range(world)]
        got = xchg_allreduce(parts)
        ref = np.zeros(shape, dtype=np.float32)
        for p in parts:                       # same fixed rank order -> identical bits
            ref = (ref + p).astype(np.float32)
        assert np.array_equal(got, ref)


def test_dino_slice_records_and_merge_equal_the_reference_loss():
    """Algorithmic identity behind dino_fwd_partial + dino_finish: per-slice log2-domain records with the
    S-trick, merged over slices, reproduce DINOLoss.forward (oracle) for ragged K and any crop count."""
    import numpy as np
    import torch
    from oracle import lafs_oracle as O
    from tests.kernel_emulation import dino_sliced_loss
    g = torch.Generator().manual_seed(0)
    for B, K, nc, tt in ((3, 1000, 6, 0.04), (2, 256, 2, 0.07), (4, 777, 10, 0.05)):
        s = torch.randn(nc * B, K, generator=g, dtype=torch.float64) * 3
        t = torch.randn(2 * B, K, generator=g, dtype=torch.float64) * 3
        c = torch.randn(1, K, generator=g, dtype=torch.float64) * 0.3
        ref = O.dino_loss(s.float(), t.float(), c.float(), nc, tt)
        loss, stats, colsum = dino_sliced_loss(s.float().numpy(), t.float().numpy(), c.float().numpy(), nc, 10.0, 1.0 / tt)
        assert abs(loss - float(ref)) <= 2e-6 * abs(float(ref)), (loss, float(ref))
        # two-level form of round 2: 8 adjacent slice records merged per CTA, then over the CTAs -- same numbers
        loss2, stats2, _ = dino_sliced_loss(s.float().numpy(), t.float().numpy(), c.float().numpy(), nc, 10.0, 1.0 / tt,
                                            slice_cols=64, cta_slices=8)
        assert abs(loss2 - loss) <= 1e-12 * abs(loss) and np.allclose(stats2, stats, rtol=1e-12, atol=1e-12)
        lse = torch.logsumexp(s.float().double() * 10.0, dim=1).view(nc, B).numpy()
        assert np.allclose(stats[:nc] * np.log(2.0), lse, rtol=1e-10, atol=1e-9)
        assert np.allclose(colsum, t.float().double().sum(0).numpy())


def test_head_chunk_and_rank_merge_equal_cross_entropy():
    """Online-softmax records per class chunk, merged per rank and then across class shards in rank order,
    give CrossEntropyLoss(CosFace logits) for any shard count (torch.chunk ownership, ragged last shard)."""
    import numpy as np
    import torch
    from oracle import lafs_oracle as O
    from tests.kernel_emulation import head_sharded_loss
    torch.manual_seed(1)
    B, C, D = 9, 1003, 32
    x, w = torch.randn(B, D), torch.randn(C, D)
    lab = torch.randint(0, C, (B,))
    lab[0], lab[1] = 0, C - 1
    cos = torch.nn.functional.normalize(x) @ torch.nn.functional.normalize(w).t()
    ref = float(O.cross_entropy(O.cosface_logits(x, w, lab), lab))
    for world in (1, 2, 4, 8):
        got = head_sharded_loss(cos.numpy(), lab.numpy(), 64.0, 0.4, world, chunk=100)
        assert abs(got - ref) <= 1e-5 * abs(ref), (world, got, ref)


def test_fused_dino_head_decomposition_equals_the_oracle():
    """(f1) the algebra of csrc/dino_head.cu (centre folded into three bf16 K columns, U = Q.W_s cross terms,
    dX from O and U, dW from one TN GEMM, weight-norm Jacobian) against autograd through the oracle on the same
    bf16-rounded operands -- checked on CPU before any GPU time is spent."""
    import torch
    from oracle import lafs_oracle as O
    from tests import kernel_emulation as KE
    g = torch.Generator().manual_seed(5)
    for (B, K, D, ncrops, gscale) in ((3, 777, 64, 6, 1.0), (5, 1500, 128, 4, 8.0), (2, 300, 64, 2, 1.0)):
        xs = torch.randn(ncrops * B, D, generator=g) * 2
        xt = torch.randn(2 * B, D, generator=g) * 2
        vs = torch.randn(K, D, generator=g) * 0.3
        vt = vs + torch.randn(K, D, generator=g) * 0.05
        gs = 0.5 + torch.rand(K, generator=g)
        gt = 0.5 + torch.rand(K, generator=g)
        center = torch.randn(K, generator=g) * 0.1
        tt, ts = 0.05, 0.1
        loss, colsum, dx, dv, dg = KE.dino_head_fused(xs, xt, vs, gs, vt, gt, center, ncrops, 1 / ts, 1 / tt, gscale)
        rl, rdx, rdv, rdg, rc1 = O.dino_head_loss_and_grads(xs, xt, vs, gs, vt, gt, center, ncrops, tt, ts,
                                                            round_bf16=True, grad_out=gscale)
        assert abs(float(loss) - float(rl)) <= 2e-4 * abs(float(rl)), (B, K, float(loss), float(rl))
        for a, b, name in ((dx, rdx, "dx"), (dv, rdv, "dv"), (dg, rdg, "dg")):
            err = float((a - b).abs().max() / b.abs().max())
            assert err < 5e-3, (name, B, K, err)
        # centre: colsum is the column sum of the teacher logits the GEMM sees
        t = O.dino_head_logits(xt, vt, gt, round_bf16=True)
        torch.testing.assert_close(colsum, t.sum(0), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(O.dino_center_update(center.reshape(1, -1), t), rc1)
