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
