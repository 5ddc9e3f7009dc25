"""GPU parity: fused landmark gather -> patch embedding (tcgen05) against golden vectors and
the oracle evaluated on bf16-rounded tokens / weights."""
import os

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


def check(out, ref, fp32):
    scale = ref.abs().max()
    err = (out.float().cpu() - ref).abs().max()
    # fp32 output: the only differences are 1-ulp bf16 rounding flips of individual token values
    tol = 1e-3 if fp32 else 1e-3 + 2 ** -8     # bf16 output rounding on top
    assert err <= tol * scale, (float(err), float(scale))


@pytest.mark.parametrize("n", [196, 36])
@pytest.mark.parametrize("dim", [768, 384, 128])
def test_gather_embed_vs_oracle(P, n, dim):
    torch.manual_seed(n + dim)
    B = 5
    imgs = torch.rand(B, 3, 112, 112) * 2 - 1
    th = torch.rand(B, n, 2) * 111 + torch.randn(B, n, 2) * 5
    th[0, 0] = torch.tensor([-30.0, 200.0]); th[0, 1] = torch.tensor([0.2, 111.7])
    lin = torch.nn.Linear(192, dim)
    ref = O.gather_embed(imgs, th, lin.weight.detach(), lin.bias.detach(), round_bf16=True)
    wts = P.PatchEmbedWeights([(lin.weight.cuda(), lin.bias.cuda())])
    (o32,) = P.gather_embed(imgs.cuda(), th.cuda(), wts, out_dtype=torch.float32)
    check(o32, ref, fp32=True)
    (o16,) = P.gather_embed(imgs.cuda(), th.cuda(), wts)
    assert o16.dtype == torch.bfloat16
    check(o16, ref, fp32=False)


def test_two_models_share_one_gather_many_faces(P):
    """student + teacher projections of the same patches; more faces than SMs (persistent loop)."""
    torch.manual_seed(3)
    B, n, dim = 333, 196, 256
    imgs = torch.rand(B, 3, 112, 112) * 2 - 1
    th = torch.rand(B, n, 2) * 111 + torch.randn(B, n, 2) * 5
    s, t = torch.nn.Linear(192, dim), torch.nn.Linear(192, dim)
    wts = P.PatchEmbedWeights([(s.weight.cuda(), s.bias.cuda()), (t.weight.cuda(), t.bias.cuda())])
    os_, ot = P.gather_embed(imgs.cuda(), th.cuda(), wts, out_dtype=torch.float32)
    for o, lin in ((os_, s), (ot, t)):
        for b in (0, 147, 148, 149, 332):
            ref = O.gather_embed(imgs[b:b + 1], th[b:b + 1], lin.weight.detach(), lin.bias.detach(), round_bf16=True)
            check(o[b:b + 1], ref, fp32=True)


def test_patch_embed_golden(P, golden):
    g, pg = golden("patch_embed"), golden("patches")
    imgs, th = T(pg["imgs"]), T(pg["theta196"])
    w, b = T(g["weight"]), T(g["bias"])           # dim 64 is not a multiple of 128: pad to 128 rows
    wp = torch.zeros(128, 192); wp[:64] = w
    bp = torch.zeros(128); bp[:64] = b
    wts = P.PatchEmbedWeights([(wp.cuda(), bp.cuda())])
    (o,) = P.gather_embed(imgs.cuda(), th.cuda(), wts, out_dtype=torch.float32)
    ref = T(g["embedded"])                        # reference fp32 patch_to_embedding output
    err = (o[:, :, :64].cpu() - ref).abs().max()
    assert err <= 1e-2 * ref.abs().max(), float(err)   # bf16 operands vs the fp32 reference


@pytest.mark.parametrize("n", [196, 36])
def test_uint8_images_normalised_in_kernel(P, n):
    """uint8 transport: ToTensor + Normalize(0.5, 0.5) fused into the gather == the reference's
    fp32 pipeline (lafs_train.py:800-803) followed by extract + patch_to_embedding."""
    torch.manual_seed(n)
    B, dim = 301, 768
    u8 = torch.randint(0, 256, (B, 3, 112, 112), dtype=torch.uint8)
    th = torch.rand(B, n, 2) * 111 + torch.randn(B, n, 2) * 5
    th[0, 0] = torch.tensor([-3.0, 113.0]); th[0, 1] = torch.tensor([400.0, 5.0])
    s, t = torch.nn.Linear(192, dim), torch.nn.Linear(192, dim)
    wts = P.PatchEmbedWeights([(s.weight.cuda(), s.bias.cuda()), (t.weight.cuda(), t.bias.cuda())])
    o_s, o_t = P.gather_embed(u8.cuda(), th.cuda(), wts, out_dtype=torch.float32)
    imgs_f = (u8.float() / 255 - 0.5) / 0.5
    for b in (0, 1, 150, 300):
        for o, lin in ((o_s, s), (o_t, t)):
            ref = O.gather_embed(imgs_f[b:b + 1], th[b:b + 1], lin.weight.detach(), lin.bias.detach(), round_bf16=True)
            check(o[b:b + 1], ref, fp32=True)
    # same numbers as the fp32-input kernel up to bf16 token rounding flips
    o_f, _ = P.gather_embed(imgs_f.cuda(), th.cuda(), wts, out_dtype=torch.float32)
    assert (o_f - o_s).abs().max() <= 1e-3 * o_f.abs().max()


@pytest.mark.parametrize("B,n,dim", [(6, 196, 768), (9, 36, 256), (150, 196, 128)])
def test_gather_embed_train_backward_vs_autograd(P, B, n, dim):
    """Training path of a3: fused forward + tcgen05 backward GEMMs against fp32 autograd through the
    oracle's differentiable gather + F.linear (bf16 operands: max-norm-relative tolerances)."""
    torch.manual_seed(B + n + dim)
    imgs = torch.rand(B, 3, 112, 112) * 2 - 1
    th = torch.rand(B, n, 2) * 100 + 5
    lin = torch.nn.Linear(192, dim)
    gout = torch.randn(B, n, dim)
    # reference: oracle gather (autograd through grid_sample) + fp32 linear
    th_r = th.clone().requires_grad_(True)
    w_r = lin.weight.detach().clone().requires_grad_(True)
    b_r = lin.bias.detach().clone().requires_grad_(True)
    tok = O.extract_tokens(imgs, th_r)
    ref = torch.nn.functional.linear(tok, w_r, b_r)
    (ref * gout).sum().backward()
    th_g = th.cuda().requires_grad_(True)
    w_g = lin.weight.detach().cuda().requires_grad_(True)
    b_g = lin.bias.detach().cuda().requires_grad_(True)
    out = P.gather_embed_train(imgs.cuda(), th_g, w_g, b_g)
    assert out.dtype == torch.bfloat16 and out.shape == (B, n, dim)
    assert (out.float().cpu() - ref.detach()).abs().max() <= (1e-3 + 2 ** -7) * ref.abs().max()
    (out.float() * gout.cuda()).sum().backward()
    for got, want, name in ((w_g.grad, w_r.grad, "weight"), (b_g.grad, b_r.grad, "bias"), (th_g.grad, th_r.grad, "theta")):
        err = (got.cpu() - want).abs().max() / want.abs().max()
        assert err <= 2e-2, (name, float(err))
        cs = torch.nn.functional.cosine_similarity(got.cpu().flatten(), want.flatten(), dim=0)
        assert cs > 0.999, (name, float(cs))


@pytest.mark.parametrize("B,n,u8", [(37, 36, True), (11, 196, True), (9, 196, False), (5, 49, True)])
def test_saved_tokens_match_oracle_tokens(P, B, n, u8):
    """gather_embed(save_tokens=): the bf16 tokens kept for the weight gradient are the oracle's tokens in the
    kernel's K order (k = c*64 + j*8 + i  <-  feature (i*8+j)*3 + c), with the ones / zero columns untouched."""
    torch.manual_seed(B + n)
    if u8:
        raw = torch.randint(0, 256, (B, 3, 112, 112), dtype=torch.uint8)
        imgs, ref_img = raw.cuda(), (raw.float() / 255 - 0.5) / 0.5
    else:
        ref_img = torch.rand(B, 3, 112, 112) * 2 - 1
        imgs = ref_img.cuda()
    th = torch.rand(B, n, 2) * 111 + torch.randn(B, n, 2) * 6
    lin = torch.nn.Linear(192, 128)
    wts = P.PatchEmbedWeights([(lin.weight.detach().cuda(), lin.bias.detach().cuda())])
    buf = P.new_token_buffer(B * n, "cuda")
    P.gather_embed(imgs, th.cuda(), wts, save_tokens=buf)
    tok = O.extract_tokens(ref_img, th).reshape(B * n, 8, 8, 3)          # [.., i, j, c]
    want = tok.permute(0, 3, 2, 1).reshape(B * n, 192)                     # k = c*64 + j*8 + i
    got = buf.float().cpu()
    assert (got[:, :192] - want).abs().max() <= 2 ** -8 * max(1.0, float(want.abs().max())) + 1e-3
    assert bool((got[:, 192] == 1).all()) and bool((got[:, 193:] == 0).all())


@pytest.mark.parametrize("B,n,dim,u8", [(9, 196, 768, True), (23, 36, 256, True), (5, 196, 128, False), (160, 36, 384, True)])
def test_sequence_epilogue_cls_pos_matches_reference_ops(P, B, n, dim, u8):
    """the fused kernel's sequence epilogue against the reference's op sequence after
    patch_to_embedding (ViT_face.py:762-768: cat(cls, x); x += pos_embedding[:, :n+1]; dropout off), two models
    (student / teacher parameters) from one gather.  Same bf16 rounding as the plain form: the sum is formed in fp32
    and rounded once."""
    torch.manual_seed(B + n + dim)
    imgs = torch.randint(0, 256, (B, 3, 112, 112), dtype=torch.uint8) if u8 else torch.rand(B, 3, 112, 112) * 2 - 1
    ref_img = (imgs.float() / 255 - 0.5) / 0.5 if u8 else imgs
    th = torch.rand(B, n, 2) * 111 + torch.randn(B, n, 2) * 5
    lins = [torch.nn.Linear(192, dim), torch.nn.Linear(192, dim)]
    pos = [torch.randn(1, 197, dim), torch.randn(1, 197, dim)]
    cls = [torch.randn(1, 1, dim), torch.randn(1, 1, dim)]
    wts = P.PatchEmbedWeights([(l.weight.detach().cuda(), l.bias.detach().cuda()) for l in lins])
    outs = P.gather_embed(imgs.cuda(), th.cuda(), wts, out_dtype=torch.float32,
                          seq=[(pos[0].cuda(), cls[0].cuda()), (pos[1].cuda(), cls[1].cuda())])
    for m in range(2):
        x = O.gather_embed(ref_img, th, lins[m].weight.detach(), lins[m].bias.detach(), round_bf16=True)
        x = torch.cat((cls[m].expand(B, -1, -1), x), dim=1)
        x = x + pos[m][:, :(n + 1)]
        assert outs[m].shape == (B, n + 1, dim)
        assert (outs[m].cpu() - x).abs().max() <= 1e-3 * x.abs().max()
        assert torch.equal(outs[m][:, 0].cpu(), (cls[m] + pos[m][:, :1]).expand(B, 1, dim)[:, 0])     # cls row: exact


def test_sequence_epilogue_dropout_is_inverted_dropout(P):
    """dropout(p) in the epilogue: every element is either 0 or the p = 0 value times 1/(1-p); the kept fraction is
    1-p; the mask depends on the seed only (reproducible), and differs between seeds."""
    torch.manual_seed(0)
    B, n, dim, p = 64, 36, 256, 0.1
    imgs = torch.randint(0, 256, (B, 3, 112, 112), dtype=torch.uint8).cuda()
    th = (torch.rand(B, n, 2) * 111).cuda()
    lin = torch.nn.Linear(192, dim)
    wts = P.PatchEmbedWeights([(lin.weight.detach().cuda(), lin.bias.detach().cuda())])
    seq = [(torch.randn(1, 197, dim).cuda(), torch.randn(1, 1, dim).cuda())]
    (base,) = P.gather_embed(imgs, th, wts, out_dtype=torch.float32, seq=seq)
    (a,) = P.gather_embed(imgs, th, wts, out_dtype=torch.float32, seq=seq, drop_p=p, seed=7)
    (a2,) = P.gather_embed(imgs, th, wts, out_dtype=torch.float32, seq=seq, drop_p=p, seed=7)
    (b,) = P.gather_embed(imgs, th, wts, out_dtype=torch.float32, seq=seq, drop_p=p, seed=8)
    assert torch.equal(a, a2) and not torch.equal(a, b)
    kept = a != 0
    torch.testing.assert_close(a[kept], (base * (1 / (1 - p)))[kept], rtol=1e-6, atol=1e-6)
    frac = float(kept.float().mean())
    assert abs(frac - (1 - p)) < 5e-3, frac
    assert abs(float((a != 0).float()[:, 0].mean()) - (1 - p)) < 2e-2                  # the cls rows are dropped too
    assert abs(float(((a != 0) & (b != 0)).float().mean()) - (1 - p) ** 2) < 5e-3       # independent masks
