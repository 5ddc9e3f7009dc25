"""CPU: host-side logic of the drop-in modules that needs no GPU -- label handling of the margin head (dense mixup
targets -> two labels + one weight, util/mixup_my.py:13-24; the opt-in range check that stands in for the
reference's one_hot.scatter_, ViT_face.py:66-68), and the deferred-logits hand-over between DINOHead and DINOLoss."""
import pytest
import torch

from oracle import lafs_oracle as O


def test_two_hot_from_dense_recovers_mixup_labels():
    from lafs_cvpr2024_b200.margin_head import two_hot_from_dense
    g = torch.Generator().manual_seed(0)
    C = 37
    lab = torch.randint(0, C, (12,), generator=g)
    lab[3] = lab[12 - 1 - 3]                                   # a row whose two classes coincide
    for lam in (0.3, 0.8, 1.0):
        tgt = O.mixup_target(lab, C, lam)                      # the oracle's restatement of mixup_target (pinned by golden)
        la, lb, w = two_hot_from_dense(tgt)
        rebuilt = torch.zeros_like(tgt)
        rebuilt.scatter_add_(1, la.view(-1, 1), torch.full((12, 1), w))
        rebuilt.scatter_add_(1, lb.view(-1, 1), torch.full((12, 1), 1.0 - w))
        same = la == lb                                        # the kernels give a coinciding pair weight 1
        rebuilt[same] = 0
        rebuilt[same, la[same]] = 1.0
        torch.testing.assert_close(rebuilt, tgt, rtol=0, atol=1e-6)
    hard = torch.nn.functional.one_hot(lab, C).float()
    la, lb, w = two_hot_from_dense(hard)
    assert torch.equal(la, lab) and torch.equal(lb, lab) and w == 1.0


def test_two_hot_from_dense_rejects_what_the_fused_head_cannot_express():
    from lafs_cvpr2024_b200.margin_head import two_hot_from_dense
    lab = torch.tensor([3, 1, 4, 1, 5, 9, 2, 6])
    per_row = O.mixup_target(lab, 10, 0.3)
    per_row[0] = O.mixup_target(lab, 10, 0.6)[0]               # Mixup 'elem' / 'pair': a different weight per row
    with pytest.raises(ValueError, match="per row"):
        two_hot_from_dense(per_row)
    three = O.mixup_target(lab, 10, 0.3)
    three[1, 7] = 0.1
    with pytest.raises(ValueError, match="more than two"):
        two_hot_from_dense(three)


def test_label_range_check_mirrors_scatter():
    import lafs_cvpr2024_b200 as P
    h = P.CosFace(16, 50, None)
    ok = torch.tensor([0, 49, 7])
    la, lb, lam = h._labels(ok, None, 1.0)
    assert la.dtype == torch.int64 and lb is None and lam == 1.0
    h.check_labels = True
    h._labels(ok, ok.flip(0), 0.4)
    for bad in (torch.tensor([0, 50]), torch.tensor([-1, 3])):
        with pytest.raises(IndexError):
            h._labels(bad, None, 1.0)
        with pytest.raises(RuntimeError):                      # what the reference does: scatter_ on the one-hot raises
            torch.zeros(2, 50).scatter_(1, bad.view(-1, 1), 1)
    with pytest.raises(ValueError):
        P.ArcFace(16, 50, None)._labels(torch.zeros(3, 50), None, 1.0)      # ArcFace takes hard labels


def test_deferred_logits_hand_over():
    import lafs_cvpr2024_b200 as P
    torch.manual_seed(1)
    hs = P.DINOHead(24, 130, nlayers=2, hidden_dim=32, bottleneck_dim=64, fused_loss=True)
    ht = P.DINOHead(24, 130, nlayers=2, hidden_dim=32, bottleneck_dim=64)
    x = torch.randn(6, 24)
    d = hs(x)
    assert isinstance(d, P.DeferredLogits) and d.shape == (6, 130) and len(d) == 6 and d.head is hs
    assert d.features.requires_grad and d.features.shape == (6, 64)
    hs.fused_loss = False
    assert torch.equal(d.logits(), hs(x))                      # the escape hatch is the reference's forward
    crit = P.DINOLoss(130, 3, 0.04, 0.07, 30, 41)
    with pytest.raises(TypeError):                             # both heads must defer, or neither
        crit(d, ht(torch.randn(4, 24)), 0)
    ht.fused_loss = True
    with pytest.raises(RuntimeError, match="no CPU fallback"):  # the fused path is CUDA-only, like every kernel path
        crit(d, ht(torch.randn(4, 24)), 0)


def test_landmark_cnn_keep_num_rule():
    """ViT_face.py:1319-1325,1363-1385: which landmarks reach the patch extractor."""
    import lafs_cvpr2024_b200 as P
    kw = dict(loss_type="None", GPU_ID=None, num_class=2, image_size=112, patch_size=8, dim=64, depth=1, heads=2,
              mlp_dim=32, stn=torch.nn.Identity())
    m196 = P.face_landmark_4simmin_glo_loc(num_patches=196, **kw)
    m144 = P.face_landmark_4simmin_glo_loc(num_patches=144, **kw)
    assert m196._keep_num(112, 196, False) == 196 and m144._keep_num(112, 144, False) == 144
    assert m196._keep_num(96, 196, False) == 144            # other input sizes: the first (H/p)^2 landmarks
    assert m196._keep_num(112, 36, True) == 36 and m196._keep_num(112, 196, True) == 196   # re-sampled sets are kept whole
