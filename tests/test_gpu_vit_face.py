"""GPU: the Part-fViT / landmark-CNN wrappers (reference module surface) against the same
composition built from the CPU oracle pieces.  The MobileNetV3 trunk is outside the hot path;
a tiny stand-in trunk with the same output contract ([B,160,h,w]) is injected."""
import copy

import pytest
import torch
import torch.nn as nn

from oracle import lafs_oracle as O

pytestmark = pytest.mark.gpu


class TinyTrunk(nn.Module):
    def __init__(self):
        super().__init__()
        self.net = nn.Sequential(nn.Conv2d(3, 16, 7, stride=8, padding=3), nn.ReLU(), nn.Conv2d(16, 160, 3, stride=4, padding=1))

    def forward(self, x):
        return self.net(x)


@pytest.fixture(scope="module")
def P():
    import lafs_cvpr2024_b200 as pkg
    return pkg


def oracle_forward(m, x, label=None):
    """The wrapper's forward with every hot-path kernel replaced by its oracle function (CPU)."""
    feat = m.stn(x).mean(dim=(-2, -1))
    theta = O.landmark_post(m.output_layer(feat))
    tok = O.extract_tokens(x, theta)
    y = m.patch_to_embedding(tok)
    b, n, _ = y.shape
    y = torch.cat((m.cls_token.expand(b, -1, -1), y), 1) + m.pos_embedding[:, :n + 1]
    emb = m.mlp_head(m.transformer(m.dropout(y))[:, 0])
    if label is None:
        return emb, theta
    return O.cross_entropy(O.cosface_logits(emb, m.loss.weight, label), label), theta


def make(P, **kw):
    torch.manual_seed(0)
    cfg = dict(loss_type="CosFace", GPU_ID=None, num_class=300, image_size=112, patch_size=8, dim=128, depth=2,
               heads=3, mlp_dim=128, num_patches=196, with_land=True, stn=TinyTrunk())
    cfg.update(kw)
    return P.ViT_face_landmark_patch8(**cfg).eval()


def test_finetune_forward_backward_matches_oracle_composition(P):
    torch.backends.cudnn.allow_tf32 = False          # the stand-in trunk must match its CPU twin closely
    torch.backends.cuda.matmul.allow_tf32 = False
    m = make(P)
    ref = copy.deepcopy(m)
    x = torch.rand(4, 3, 112, 112) * 2 - 1
    lab = torch.tensor([0, 7, 299, 150])
    loss_ref, theta_ref = oracle_forward(ref, x, lab)
    loss_ref.backward()
    mg = m.cuda()
    logits, theta = mg(x.cuda(), lab.cuda())
    assert logits.shape == (4, 300)
    # the stand-in trunk runs through cuDNN here and through CPU convs in the oracle composition, so
    # `raw` differs in its last bits; the tail itself is bit-exact (test_gpu_patches.py)
    torch.testing.assert_close(theta.detach().cpu(), theta_ref.detach(), rtol=0, atol=2e-2)
    loss = torch.nn.CrossEntropyLoss()(logits, lab.cuda())
    loss.backward()
    assert abs(float(loss) - float(loss_ref)) <= 5e-3 * abs(float(loss_ref))
    for name in ("patch_to_embedding.weight", "output_layer.1.weight", "stn.net.0.weight", "loss.weight"):
        g = dict(mg.named_parameters())[name].grad.cpu()
        gr = dict(ref.named_parameters())[name].grad
        err = (g - gr).abs()
        assert err.max() <= 3e-2 * gr.abs().max(), (name, float(err.max()), float(gr.abs().max()),
                                                    int(err.flatten().argmax()), tuple(g.shape))
    # fused loss entry point gives the same loss
    mg.zero_grad()
    l2 = mg.forward_loss(x.cuda(), lab.cuda())
    assert abs(float(l2) - float(loss)) <= 1e-4 * abs(float(loss))


def test_inference_uses_fused_path(P):
    m = make(P, loss_type="None")
    ref = copy.deepcopy(m)
    x = torch.rand(3, 3, 112, 112) * 2 - 1
    with torch.no_grad():
        emb_ref, _ = oracle_forward(ref, x)
        emb = m.cuda()(x.cuda())
    assert (emb.cpu() - emb_ref).abs().max() <= 2e-2 * emb_ref.abs().max()     # bf16 patch embedding inside


def test_full_precision_model_keeps_fp32_patch_embedding(P):
    """fp16=False and no autocast region: patch_to_embedding stays an fp32 nn.Linear on the (bit-exact) gathered tokens,
    with and without gradients; inside an autocast region the same model takes the tensor-core path again."""
    torch.backends.cudnn.allow_tf32 = False          # the stand-in trunk must match its CPU twin closely
    torch.backends.cuda.matmul.allow_tf32 = False
    m = make(P, loss_type="None", fp16=False)
    ref = copy.deepcopy(m)
    x = torch.rand(3, 3, 112, 112) * 2 - 1
    with torch.no_grad():
        emb_ref, _ = oracle_forward(ref, x)
        emb = m.cuda()(x.cuda())
    # fp32 everywhere: 10x tighter than the bf16 path's 2e-2 (what is left is the cuDNN / CPU difference of the trunk)
    assert (emb.cpu() - emb_ref).abs().max() <= 2e-3 * emb_ref.abs().max()
    proj = torch.randn(emb_ref.shape[1], generator=torch.Generator().manual_seed(2))   # the output is LayerNorm'ed:
    e = m(x.cuda())                                   # eval mode (the landmark head has a Dropout), gradients on
    (e @ proj.cuda()).sum().backward()                # a random projection has a real gradient, e.square().mean() has none
    er, _ = oracle_forward(ref, x)
    (er @ proj).sum().backward()
    assert (e.detach().cpu() - er.detach()).abs().max() <= 2e-3 * er.detach().abs().max()
    gw, gr = m.patch_to_embedding.weight.grad.cpu(), ref.patch_to_embedding.weight.grad
    assert (gw - gr).abs().max() <= 1e-2 * gr.abs().max()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        emb_ac = m(x.cuda())
    assert (emb_ac.float().cpu() - emb_ref).abs().max() <= 5e-2 * emb_ref.abs().max()


def test_landmark_cnn_wrapper_ssl_calls(P):
    torch.manual_seed(1)
    cnn = P.face_landmark_4simmin_glo_loc(loss_type="CosFace", GPU_ID=None, num_class=10, num_patches=196, image_size=112,
                                          patch_size=8, dim=64, depth=1, heads=2, mlp_dim=64, stn=TinyTrunk()).eval()
    ref = copy.deepcopy(cnn)
    x = torch.rand(3, 3, 112, 112) * 2 - 1
    xa = torch.rand(3, 3, 112, 112) * 2 - 1
    g = cnn.cuda()
    ps = torch.tensor([8, 8])
    with torch.no_grad():
        raw = g.output_layer(g.stn(x.cuda()).mean(dim=(-2, -1))).cpu()   # same trunk output as the module sees
        # global view (lafs_train.py:535): noise only
        torch.manual_seed(5)
        th, mos = g(x.cuda(), x_Aug=xa.cuda(), patch_shape=ps, Random_prob=True, return_prob=True)
        torch.manual_seed(5)
        th_ref = O.landmark_post(raw, torch.randn(3, 196, 2) * 5)
        assert torch.equal(th.cpu(), th_ref) and torch.equal(mos.cpu(), O.extract_patches(xa, th_ref, 196))
        # local view (lafs_train.py:565): noise + 36 re-sampled landmarks
        torch.manual_seed(6)
        th, mos = g(x.cuda(), x_Aug=xa.cuda(), patch_shape=ps, Random_prob=True, ran_sample=True)
        torch.manual_seed(6)
        noise = torch.randn(3, 196, 2) * 5
        idx = torch.randint(0, 196, (3, 36, 1))
        th_ref = O.landmark_post(raw, noise, idx)
        assert mos.shape == (3, 3, 48, 48)
        assert torch.equal(th.cpu(), th_ref) and torch.equal(mos.cpu(), O.extract_patches(xa, th_ref, 36))


def test_default_landmark_trunk_runs_self_contained():
    """No reference checkout, no injected stn: the package's own MobileNetV3-large trunk
    (landmark_trunk.py, reference checkpoint keys) feeds the landmark tail and the gather kernels."""
    import lafs_cvpr2024_b200 as P
    torch.manual_seed(0)
    g = P.face_landmark_4simmin_glo_loc(loss_type="None", GPU_ID=None, num_class=0, image_size=112, patch_size=8,
                                        dim=64, depth=1, heads=2, mlp_dim=64).cuda().eval()
    assert any(k.startswith("stn.features.15.conv.8.") for k in g.state_dict())
    x = torch.rand(4, 3, 112, 112, device="cuda") * 2 - 1
    with torch.no_grad():
        theta, patches = g(x, Random_prob=True, return_prob=True)
    assert theta.shape == (4, 196, 2) and patches.shape == (4, 3, 112, 112)
    assert torch.isfinite(theta).all() and torch.isfinite(patches).all()
    m = P.ViT_face_landmark_patch8(loss_type="CosFace", GPU_ID=None, num_class=100, image_size=112, patch_size=8,
                                   dim=128, depth=1, heads=2, mlp_dim=128, num_patches=196, with_land=True).cuda().train()
    lab = torch.randint(0, 100, (4,), device="cuda")
    loss = m.forward_loss(x, lab)                      # fused gather->embed forward, tcgen05 backward GEMMs
    loss.backward()
    assert torch.isfinite(loss)
    for name in ("patch_to_embedding.weight", "output_layer.1.weight", "stn.features.0.0.weight", "loss.weight"):
        gr = dict(m.named_parameters())[name].grad
        assert gr is not None and torch.isfinite(gr).all() and float(gr.abs().max()) > 0, name
