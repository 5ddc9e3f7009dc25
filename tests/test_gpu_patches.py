"""GPU parity: landmark tail and bilinear patch gather against golden vectors and the oracle."""
import numpy as np
import pytest
import torch

from oracle import lafs_oracle as O

pytestmark = pytest.mark.gpu


def T(a):
    return torch.from_numpy(np.asarray(a))


@pytest.fixture(scope="module")
def P():
    import lafs_cvpr2024_b200 as pkg
    return pkg


PS = torch.tensor([8, 8])


def test_gather_golden_bit_exact(P, golden):
    g = golden("patches")
    imgs = T(g["imgs"]).cuda()
    for n in (196, 36):
        out = P.extract_patches_pytorch_gridsample(imgs, T(g[f"theta{n}"]).cuda(), PS, n)
        ref = T(g[f"mosaic{n}"])
        assert out.shape == ref.shape
        d = (out.cpu() - ref).abs().max()
        assert d <= 1e-5, d                      # north-star tolerance for fp32 patch samples
        assert torch.equal(out.cpu(), ref), d    # the kernel reproduces the op order exactly
        tok = P.extract_tokens(imgs, T(g[f"theta{n}"]).cuda())
        assert torch.equal(tok.cpu(), O.tokens_from_mosaic(ref))


def test_gather_random_vs_oracle_both_coord_modes(P):
    from lafs_cvpr2024_b200 import patches, _lib
    torch.manual_seed(0)
    B = 16
    imgs = torch.rand(B, 3, 112, 112) * 2 - 1
    th = torch.rand(B, 196, 2) * 111 + torch.randn(B, 196, 2) * 5
    th[0, :4] = torch.tensor([[1e6, 3.0], [-1e6, 3.0], [50.0, 1e9], [55.5, -7.5]])  # far outside
    out = P.extract_patches_pytorch_gridsample(imgs.cuda(), th.cuda(), PS, 196).cpu()
    assert torch.equal(out, O.extract_patches(imgs, th, 196))
    # reciprocal-multiply mode == what the reference does in eager CUDA (oracle ops run on the GPU)
    old = patches.COORD_MODE
    try:
        patches.COORD_MODE = _lib.COORD_RECIP
        out_r = P.extract_patches_pytorch_gridsample(imgs.cuda(), th.cuda(), PS, 196)
    finally:
        patches.COORD_MODE = old
    eager = O.extract_patches(imgs.cuda(), th.cuda(), 196)
    assert (out_r - eager).abs().max() <= 1e-5
    assert torch.equal(out_r.cpu(), O.extract_patches(imgs, th, 196, recip_mul=True))


def test_gather_other_shapes_and_empty(P):
    torch.manual_seed(1)
    for (B, C, H, n) in [(1, 3, 112, 1), (2, 1, 64, 4), (3, 3, 96, 49), (2, 2, 112, 144)]:
        imgs = torch.randn(B, C, H, H)
        th = torch.rand(B, n, 2) * (H - 1)
        out = P.extract_patches_pytorch_gridsample(imgs.cuda(), th.cuda(), PS, n).cpu()
        assert torch.equal(out, O.extract_patches(imgs, th, n)), (B, C, H, n)
    out = P.extract_patches_pytorch_gridsample(torch.zeros(0, 3, 112, 112).cuda(), torch.zeros(0, 196, 2).cuda(), PS, 196)
    assert out.shape == (0, 3, 112, 112)
    with pytest.raises(ValueError):
        P.extract_patches_pytorch_gridsample(torch.zeros(1, 3, 112, 112).cuda(), torch.zeros(1, 196, 2).cuda(),
                                             torch.tensor([10, 10]), 196)


def test_gather_backward_vs_oracle_autograd(P):
    torch.manual_seed(2)
    B, n = 3, 36
    imgs = (torch.rand(B, 3, 112, 112) * 2 - 1).requires_grad_(True)
    th = (torch.rand(B, n, 2) * 100 + 5.3).requires_grad_(True)
    w = torch.randn(B, 3, 48, 48)
    (O.extract_patches(imgs, th, n) * w).sum().backward()
    ig = imgs.detach().cuda().requires_grad_(True)
    tg = th.detach().cuda().requires_grad_(True)
    (P.extract_patches_pytorch_gridsample(ig, tg, PS, n) * w.cuda()).sum().backward()
    assert (tg.grad.cpu() - th.grad).abs().max() <= 1e-4 * th.grad.abs().max()
    assert (ig.grad.cpu() - imgs.grad).abs().max() <= 1e-5 * imgs.grad.abs().max() + 1e-6
    # token layout backward
    imgs.grad = None; th.grad = None
    w2 = torch.randn(B, n, 192)
    (O.extract_tokens(imgs, th) * w2).sum().backward()
    ig.grad = None; tg.grad = None
    (P.extract_tokens(ig, tg) * w2.cuda()).sum().backward()
    assert (tg.grad.cpu() - th.grad).abs().max() <= 1e-4 * th.grad.abs().max()
    assert (ig.grad.cpu() - imgs.grad).abs().max() <= 1e-5 * imgs.grad.abs().max() + 1e-6


def test_landmark_post_golden_and_rng_order(P, golden):
    g = golden("landmark_post")
    # same CPU-generator call order as the reference: randn(theta.shape)*5, then randint
    torch.manual_seed(int(g["seed_global"]))
    noise = torch.randn(3, 196, 2) * 5
    th = P.landmark_post(T(g["raw_global"]).cuda(), noise.cuda())
    assert torch.equal(th.cpu(), T(g["theta_global"]))
    torch.manual_seed(int(g["seed_local"]))
    noise = torch.randn(3, 196, 2) * 5
    idx = torch.randint(0, 196, (3, 36, 1))
    th = P.landmark_post(T(g["raw_local"]).cuda(), noise.cuda(), idx.cuda())
    assert torch.equal(th.cpu(), T(g["theta_local"]))            # bit-exact landmark indices + values
    th = P.landmark_post(T(g["raw_local"]).cuda())
    assert torch.equal(th.cpu(), T(g["theta_plain"]))
    mos = P.extract_patches_pytorch_gridsample(T(g["x_aug"]).float().cuda(), T(g["theta_local"]).cuda(), PS, 36)
    assert torch.equal(mos.cpu(), T(g["mosaic_local_from_half"]))


def test_landmark_post_backward(P):
    torch.manual_seed(4)
    raw = torch.randn(5, 392, requires_grad=True)
    w = torch.randn(5, 196, 2)
    (O.landmark_post(raw) * w).sum().backward()
    rg = raw.detach().cuda().requires_grad_(True)
    (P.landmark_post(rg) * w.cuda()).sum().backward()
    assert (rg.grad.cpu() - raw.grad).abs().max() <= 1e-4 * raw.grad.abs().max()
