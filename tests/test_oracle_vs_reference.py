"""Build container only (needs /root/reference): the oracle restatement against the LIVE
reference on fresh random inputs -- complements the committed golden vectors."""
import numpy as np
import pytest
import torch

from oracle import lafs_oracle as O

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_harness
    return ref_harness.load()


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_extract_patches_bit_exact(ref, seed):
    g = torch.Generator().manual_seed(seed)
    B = 3
    imgs = torch.randn(B, 3, 112, 112, generator=g)
    for n in (196, 36, 49):
        th = torch.rand(B, n, 2, generator=g) * 111 + torch.randn(B, n, 2, generator=g) * 5
        out = ref.VF.extract_patches_pytorch_gridsample(imgs, th, torch.tensor([8, 8]), n)
        assert torch.equal(out, O.extract_patches(imgs, th, n))


@pytest.mark.parametrize("seed", [0, 1])
def test_dino_loss_matches(ref, seed):
    g = torch.Generator().manual_seed(seed)
    B, K, nc = 5, 2048, 10
    dl = ref.L.DINOLoss(K, nc, 0.04, 0.07, 30, 41)
    dl.center = torch.randn(1, K, generator=g) * 0.2
    c0 = dl.center.clone()
    s = torch.randn(nc * B, K, generator=g) * 2
    t = torch.randn(2 * B, K, generator=g) * 2
    for epoch in (0, 12, 40):
        dl.center = c0.clone()
        loss = dl(s, t, epoch)
        temp = float(O.teacher_temp_schedule(0.04, 0.07, 30, 41)[epoch])
        assert torch.equal(loss, O.dino_loss(s, t, c0, nc, temp))
        assert torch.equal(dl.center, O.dino_center_update(c0, t))


def test_cosface_and_mixup_match(ref):
    import contextlib
    import io
    torch.manual_seed(3)
    B, D, C = 16, 48, 333
    with contextlib.redirect_stdout(io.StringIO()):
        head = ref.VF.CosFace(D, C, None)
    x = torch.randn(B, D)
    lab = torch.randint(0, C, (B,))
    assert torch.equal(head(x, lab), O.cosface_logits(x, head.weight.detach(), lab))
    soft = ref.mixup.mixup_target(lab, C, lam=0.7, smoothing=0.0, device="cpu")
    assert torch.equal(soft, O.mixup_target(lab, C, 0.7))
    assert torch.equal(head(x, soft), O.cosface_logits(x, head.weight.detach(), soft))


def test_cosine_scheduler_matches(ref):
    a = ref.dutils.cosine_scheduler(0.996, 1, 41, 123)
    assert np.array_equal(a, O.cosine_scheduler(0.996, 1, 41, 123))


def test_module_wrappers_keep_state_dict_keys(ref):
    """Checkpoint contract: the drop-in modules expose the reference's parameter names
    and shapes, so the reference's checkpoints load with strict=True."""
    import contextlib
    import io
    import lafs_cvpr2024_b200 as P
    kw = dict(loss_type="CosFace", GPU_ID=None, num_class=50, image_size=112, patch_size=8, dim=64, depth=2,
              heads=3, mlp_dim=96, num_patches=196)
    with contextlib.redirect_stdout(io.StringIO()):
        r1 = ref.VF.ViT_face_landmark_patch8(with_land=True, **kw)
        r2 = ref.VF.face_landmark_4simmin_glo_loc(**kw)
        m1 = P.ViT_face_landmark_patch8(with_land=True, **kw)
        m2 = P.face_landmark_4simmin_glo_loc(**kw)
    for r, m in ((r1, m1), (r2, m2)):
        rs, ms = r.state_dict(), m.state_dict()
        assert set(rs) == set(ms), (sorted(set(rs) ^ set(ms))[:10])
        assert all(rs[k].shape == ms[k].shape for k in rs)
        m.load_state_dict(rs, strict=True)


def test_landmark_trunk_matches_reference(ref):
    """The restated MobileNetV3-large trunk (landmark_trunk.py, outside the hot path) has the
    reference's checkpoint keys and computes the same function from the same weights."""
    import contextlib
    import io
    from face_pre_pro.mobilenet import MobileNetV3_backbone
    from lafs_cvpr2024_b200.landmark_trunk import MobileNetV3LargeTrunk
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        r = MobileNetV3_backbone(mode="large")
    m = MobileNetV3LargeTrunk()
    rs, ms = r.state_dict(), m.state_dict()
    assert list(rs) == list(ms)
    assert all(rs[k].shape == ms[k].shape for k in rs)
    m.load_state_dict(rs, strict=True)
    x = torch.randn(2, 3, 112, 112)
    for mode in ("eval", "train"):
        getattr(r, mode)(); getattr(m, mode)()
        torch.manual_seed(1); a = r(x)
        torch.manual_seed(1); b = m(x)
        assert a.shape == (2, 160, 4, 4)
        assert torch.equal(a, b)
    # same initialisation statistics (mobilenet.py:315-328): conv std = sqrt(2 / fan_out), BN (1, 0)
    m2 = MobileNetV3LargeTrunk()
    w = m2.features[0][0].weight
    assert abs(float(w.std()) - (2.0 / (16 * 9)) ** 0.5) < 0.03
    assert float(m2.features[1].conv[1].weight.min()) == 1.0 and float(m2.features[1].conv[1].bias.abs().max()) == 0.0


@pytest.mark.parametrize("random_prob,shuffle", [(False, False), (True, False), (True, True)])
def test_standard_grid_and_plain_grid_paths_match_reference(ref, random_prob, shuffle):
    """ViT_face_landmark_patch8 on image input without the landmark CNN: `use_standcoord` (fixed-grid patches,
    optional jitter / re-sampling drawn on the CPU generator in the reference's order, transposed mosaic) and the
    plain ViT patch grid.  The wrapper's re-layout logic is run here on CPU with the ORACLE's patch extractor in
    place of the gather kernel and compared with the unmodified reference module (same weights, same seed)."""
    import contextlib
    import io
    from oracle import lafs_oracle as O
    from lafs_cvpr2024_b200 import vit_face as V
    kw = dict(loss_type="None", GPU_ID=None, num_class=0, image_size=112, patch_size=8, dim=32, depth=1, heads=2,
              mlp_dim=48, num_patches=196)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        r = ref.VF.ViT_face_landmark_patch8(use_standcoord=True, Random_prob=random_prob, shuffle=shuffle, **kw).eval()
        m = V.ViT_face_landmark_patch8(use_standcoord=True, Random_prob=random_prob, shuffle=shuffle, **kw).eval()
    m.load_state_dict(r.state_dict(), strict=True)
    x = torch.rand(3, 3, 112, 112) * 2 - 1
    torch.manual_seed(5)
    with torch.no_grad():
        want = r(x)
    # the wrapper's composition with the oracle extractor (the product path substitutes the CUDA gather kernel)
    torch.manual_seed(5)
    theta, tok = V.standard_grid_tokens(x, 196, random_prob, shuffle, lambda im, th: O.extract_patches(im, th, th.shape[1]))
    with torch.no_grad():
        emb = m(tok)                                   # 3-D token input: patch_to_embedding + transformer
    assert theta.shape == (3, 196, 2)
    torch.testing.assert_close(emb, want, rtol=1e-5, atol=1e-6)
    # plain grid (no landmarks, no standard coordinates): pure PyTorch, runs on CPU through the wrapper itself
    with contextlib.redirect_stdout(io.StringIO()):
        r2 = ref.VF.ViT_face_landmark_patch8(**kw).eval()
        m2 = V.ViT_face_landmark_patch8(**kw).eval()
    m2.load_state_dict(r2.state_dict(), strict=True)
    with torch.no_grad():
        torch.testing.assert_close(m2(x), r2(x), rtol=1e-5, atol=1e-6)


def test_attention_with_mask_matches_reference(ref):
    """Masked attention (ViT_face.py:140-182): masked logits are replaced by -finfo.max, a fully masked query
    row attends uniformly (finite), and `attention_score` is stashed -- same weights, same output (CPU)."""
    from lafs_cvpr2024_b200 import vit_face as V
    torch.manual_seed(3)
    dim, heads, n = 64, 3, 9
    ra = ref.VF.Attention(dim, heads=heads, dim_head=16)
    oa = V.Attention(dim, heads=heads, dim_head=16)
    oa.load_state_dict(ra.state_dict(), strict=True)
    # the reference broadcasts its [b,n,n] mask against [b,h,n,n] logits without a head axis (ViT_face.py:171-172),
    # which only works for b == 1 (or b == heads, where it silently masks per HEAD): compare one sample at a time
    for masked_all in (False, True):
        x = torch.randn(1, n + 1, dim)
        mask = torch.rand(1, n) > 0.4
        if masked_all:
            mask[0, :] = False              # every patch token masked: their query rows are fully masked
        out_r = ra(x, mask=mask.clone())
        out_o = oa(x, mask=mask.clone())
        assert torch.isfinite(out_o).all()
        torch.testing.assert_close(out_o, out_r, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(oa.attention_score, ra.attention_score, rtol=1e-5, atol=1e-7)
    x = torch.randn(2, n + 1, dim)
    torch.testing.assert_close(oa(x), ra(x), rtol=1e-5, atol=1e-6)     # unmasked path (SDPA) agrees too
    # ... and `attention_score` of the unmasked forward (read by util/utils.py:662) is what the reference stashed
    assert oa.attention_score.shape == (2, heads, n + 1, n + 1) and not oa.attention_score.requires_grad
    torch.testing.assert_close(oa.attention_score, ra.attention_score, rtol=1e-5, atol=1e-7)
    assert V.Attention(dim, heads=heads, dim_head=16).attention_score == 0          # before any forward (ViT_face.py:159)
    assert torch.isfinite(oa(x, mask=torch.zeros(2, n, dtype=torch.bool))).all()   # batched masks work here


def test_clip_gradients_matches_reference(ref):
    """oracle clip_gradients_ (a list of gradients) == utils.clip_gradients (a model), norms and clipped gradients."""
    torch.manual_seed(5)
    model = torch.nn.Sequential(torch.nn.Linear(40, 30), torch.nn.LayerNorm(30), torch.nn.Linear(30, 7))
    for i, p in enumerate(model.parameters()):
        p.grad = torch.randn_like(p) * (10.0 if i % 2 == 0 else 0.01)
    list(model.parameters())[3].grad = None
    grads = [None if p.grad is None else p.grad.clone() for p in model.parameters()]
    norms_ref = ref.dutils.clip_gradients(model, 3.0)
    norms = O.clip_gradients_(grads, 3.0)
    assert norms == norms_ref
    for p, g in zip(model.parameters(), grads):
        assert (p.grad is None) == (g is None)
        if g is not None:
            assert torch.equal(p.grad, g)


def test_dino_head_tail_matches_reference(ref):
    """(f1) the live reference DINOHead (vision_transformer.py:265-301) + DINOLoss against the oracle restatement,
    and our DINOHead module: same state-dict keys, same init statistics, reference checkpoints load strict."""
    import warnings
    torch.manual_seed(11)
    in_dim, K, B, nc = 40, 900, 3, 5
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        hs = ref.vt.DINOHead(in_dim, K, nlayers=2, hidden_dim=64, bottleneck_dim=64, norm_last_layer=False)
        ht = ref.vt.DINOHead(in_dim, K, nlayers=2, hidden_dim=64, bottleneck_dim=64)
    with torch.no_grad():
        hs.last_layer.weight_g.uniform_(0.5, 1.5)
    fs, ft = torch.randn(nc * B, in_dim), torch.randn(2 * B, in_dim)
    dl = ref.L.DINOLoss(K, nc, 0.04, 0.07, 30, 41)
    dl.center = torch.randn(1, K) * 0.1
    c0 = dl.center.clone()
    xs = hs.mlp(fs).detach().requires_grad_(True)
    s_out = hs.last_layer(torch.nn.functional.normalize(xs, dim=-1, p=2))
    assert torch.equal(s_out, hs(fs))
    with torch.no_grad():
        xt = ht.mlp(ft)
        t_out = ht(ft)
    loss = dl(s_out, t_out, 2)
    loss.backward()
    temp = float(O.teacher_temp_schedule(0.04, 0.07, 30, 41)[2])
    ol, odx, odv, odg, oc1 = O.dino_head_loss_and_grads(xs, xt, hs.last_layer.weight_v, hs.last_layer.weight_g,
                                                        ht.last_layer.weight_v, ht.last_layer.weight_g, c0, nc, temp)
    torch.testing.assert_close(ol, loss.detach(), rtol=1e-6, atol=0)
    for a, b in ((odx, xs.grad), (odv, hs.last_layer.weight_v.grad), (odg.reshape(-1, 1), hs.last_layer.weight_g.grad)):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5 * float(b.abs().max()))
    torch.testing.assert_close(oc1, dl.center, rtol=1e-6, atol=1e-7)
    # module surface
    import lafs_cvpr2024_b200 as P
    mine = P.DINOHead(in_dim, K, nlayers=2, hidden_dim=64, bottleneck_dim=64, norm_last_layer=False)
    assert list(mine.state_dict().keys()) == list(hs.state_dict().keys())
    mine.load_state_dict(hs.state_dict(), strict=True)
    assert torch.equal(mine(fs), hs(fs))                       # default mode = the reference's forward
    assert mine.last_layer.weight_g.requires_grad and not P.DINOHead(8, 16).last_layer.weight_g.requires_grad
    mine.fused_loss = True
    d = mine(fs)
    assert isinstance(d, P.DeferredLogits) and torch.equal(d.features, hs.mlp(fs)) and torch.equal(d.logits(), hs(fs))
    big = P.DINOHead(64, 32, nlayers=3, hidden_dim=512, bottleneck_dim=256)
    assert abs(float(big.mlp[0].weight.std()) - 0.02) < 0.002 and float(big.mlp[0].bias.abs().max()) == 0.0
    assert float(big.last_layer.weight_g.min()) == 1.0 == float(big.last_layer.weight_g.max())
