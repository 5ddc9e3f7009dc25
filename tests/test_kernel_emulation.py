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
