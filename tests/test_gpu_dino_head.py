"""GPU parity of SURVEY 8f row 1: DINOHead.last_layer (weight-normed Linear on F.normalize'd features,
vision_transformer.py:284,296-300) fused with DINOLoss (lafs_train.py:643-679) -- loss, centre update and the
gradients w.r.t. the bottleneck features and last_layer.weight_v / weight_g, through the package -> C ABI ->
sm_100a kernels, against golden vectors produced by the unmodified reference (tests/golden/dino_head.npz) and the
oracle."""
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
    from lafs_cvpr2024_b200 import _lib
    assert _lib.lib().lafs_device_ok() == 1, "tests must run on a compute-capability 10.x device"
    return pkg


def maxrel(a, b):
    return float((a.float().cpu() - b.float().cpu()).abs().max() / b.float().abs().max())


def assert_colsum(colsum, xt, vt, gt):
    """column sums of the teacher logits (the input of update_center, lafs_train.py:674) against the oracle on the
    same bf16-rounded operands.  The kernel's 1/||v|| can differ from torch's in the last fp32 bit (summation order),
    which flips a bf16 rounding of an operand element here and there: a flipped element moves ONE column sum by at most
    ulp_bf16(|w|max) * |xsum|max -- the hard bound; all but a handful of columns agree to fp32 accuracy."""
    xt, vt, gt = xt.float().cpu(), vt.float().cpu(), gt.float().cpu().reshape(-1)
    t = O.dino_head_logits(xt, vt, gt, round_bf16=True)
    ref = t.sum(0)
    w = vt * (gt / vt.norm(dim=1)).unsqueeze(1)
    xsum = torch.nn.functional.normalize(xt, dim=-1).bfloat16().float().sum(0)
    err = (colsum.float().cpu() - ref).abs()
    scale = float(ref.abs().max())
    assert float(err.max()) <= 2.0 ** -7 * float(w.abs().max()) * float(xsum.abs().max()) + 1e-4 * scale
    assert float((err > 1e-4 * scale).float().mean()) < 2e-3


def run_fused(P, xs, xt, vs, gs, vt, gt, center, ncrops, ts, tt, grad_out=1.0):
    dev = "cuda"
    loss, colsum, saved = P.dino_head_forward(xs.to(dev), xt.to(dev), vs.to(dev), gs.to(dev), vt.to(dev), gt.to(dev),
                                              center.to(dev), ncrops, 1.0 / ts, 1.0 / tt)
    g = torch.tensor(float(grad_out), device=dev)
    dx, dv, dg = P.dino_head_backward(saved, g, want_grad_g=True)
    torch.cuda.synchronize()
    return loss.cpu(), colsum.cpu(), dx.cpu(), dv.cpu(), dg.cpu()


def test_fused_dino_head_golden(P, golden):
    """inputs and outputs of the UNMODIFIED reference (student/teacher DINOHead tails + DINOLoss on CPU, fp32)."""
    g = golden("dino_head")
    ncrops, temp = int(g["ncrops"]), float(g["temp"])
    xs, xt = T(g["xs"]), T(g["xt"])
    loss, colsum, dx, dv, dg = run_fused(P, xs, xt, T(g["vs"]), T(g["gs"]), T(g["vt"]), T(g["gt"]), T(g["center0"]),
                                         ncrops, 0.1, temp)
    ref = float(g["loss"])
    assert abs(float(loss) - ref) <= 1e-3 * abs(ref), (float(loss), ref)      # north star: 1e-3 relative
    # gradients against the fp32 reference: the bf16 operand rounding alone is a few 1e-3 (tests/test_kernel_emulation.py)
    assert maxrel(dx, T(g["grad_xs"])) < 2e-2
    assert maxrel(dv, T(g["grad_vs"])) < 2e-2
    assert maxrel(dg, T(g["grad_gs"]).reshape(-1)) < 2e-2
    # centre: the reference's update from its own fp32 teacher logits
    c1 = O.dino_center_update(T(g["center0"]), torch.zeros(xt.shape[0], 1), allreduced_sum=colsum.reshape(1, -1))
    torch.testing.assert_close(c1, T(g["center1"]), rtol=0, atol=2e-3 * float(T(g["center1"]).abs().max()) + 1e-4)


@pytest.mark.parametrize("B,K,D,ncrops,dtype,unit_g,grad_out", [
    (4, 1000, 64, 6, torch.float32, True, 1.0),
    (16, 4099, 256, 6, torch.float32, False, 1.0),
    (130, 3000, 128, 4, torch.float32, True, 1024.0),      # several M tiles, ragged rows, AMP-like loss scale
    (8, 777, 64, 2, torch.float16, False, 1.0),            # two global crops only; half-precision features
    (32, 65536, 256, 6, torch.bfloat16, True, 1.0),        # the reference's out_dim and bottleneck width
    (3, 520, 704, 3, torch.float32, True, 1.0),            # widest supported bottleneck (704 + 64 centre columns)
])
def test_fused_dino_head_vs_oracle_on_same_bf16_operands(P, B, K, D, ncrops, dtype, unit_g, grad_out):
    g = torch.Generator().manual_seed(B * 7 + K)
    xs = (torch.randn(ncrops * B, D, generator=g) * 2).to(dtype)
    xt = (torch.randn(2 * B, D, generator=g) * 2).to(dtype)
    vs = torch.randn(K, D, generator=g) * 0.3
    vt = vs + torch.randn(K, D, generator=g) * 0.05
    gs = torch.ones(K) if unit_g else 0.5 + torch.rand(K, generator=g)
    gt = torch.ones(K) if unit_g else 0.5 + torch.rand(K, generator=g)
    center = torch.randn(K, generator=g) * 0.1
    ts, tt = 0.1, 0.04
    loss, colsum, dx, dv, dg = run_fused(P, xs, xt, vs, gs, vt, gt, center, ncrops, ts, tt, grad_out)
    rl, rdx, rdv, rdg, rc1 = O.dino_head_loss_and_grads(xs.float(), xt.float(), vs, gs, vt, gt, center, ncrops, tt, ts,
                                                        round_bf16=True, grad_out=grad_out)
    assert abs(float(loss) - float(rl)) <= 1e-3 * abs(float(rl)), (float(loss), float(rl))
    assert maxrel(dx, rdx) < 5e-3
    assert maxrel(dv, rdv) < 5e-3
    # d/d weight_g = <v_hat, dW>: the component of dW the weight-norm Jacobian removes, a near-cancelling sum of the
    # bf16 probabilities (measured on B200: 5.1e-3 at K = 4099; the CPU emulation of the same rounding gives 2-5e-3)
    assert maxrel(dg, rdg) < 1e-2
    assert_colsum(colsum, xt, vt, gt)


def test_fused_dino_head_extreme_centre_and_temperature(P):
    """a centre far from the logits and the sharpest teacher temperature: (t - c)/0.04 spans +-60; the three-term
    bf16 split must carry the fp32 centre exactly and the online softmax must not overflow."""
    g = torch.Generator().manual_seed(11)
    B, K, D, ncrops = 8, 2048, 64, 6
    xs = torch.randn(ncrops * B, D, generator=g)
    xt = torch.randn(2 * B, D, generator=g)
    vs = torch.randn(K, D, generator=g)
    vt = torch.randn(K, D, generator=g)
    one = torch.ones(K)
    center = torch.randn(K, generator=g) * 1.4 + 0.123456789
    loss, colsum, dx, dv, _ = run_fused(P, xs, xt, vs, one, vt, one, center, ncrops, 0.1, 0.04)
    rl, rdx, rdv, _, _ = O.dino_head_loss_and_grads(xs, xt, vs, one, vt, one, center, ncrops, 0.04, 0.1, round_bf16=True)
    assert torch.isfinite(loss)
    assert abs(float(loss) - float(rl)) <= 1e-3 * abs(float(rl)), (float(loss), float(rl))
    assert maxrel(dx, rdx) < 5e-3 and maxrel(dv, rdv) < 5e-3


def test_fused_dino_head_baseline_size_against_fp32_on_device(P):
    """BASELINE configs[1] size (B=256, out_dim 65536, 2+4 crops, bottleneck 256): the oracle restatement itself,
    evaluated in fp32 on the GPU (no TF32) on the same bf16-rounded operands, plus run-to-run determinism."""
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator().manual_seed(3)
    B, K, D, ncrops = 256, 65536, 256, 6
    dev = "cuda"
    xs = (torch.randn(ncrops * B, D, generator=g) * 2).to(dev)
    xt = (torch.randn(2 * B, D, generator=g) * 2).to(dev)
    vs = (torch.randn(K, D, generator=g) * 0.02).to(dev)
    vt = (vs.cpu() + torch.randn(K, D, generator=g) * 0.002).to(dev)
    one = torch.ones(K, device=dev)
    center = (torch.randn(K, generator=g) * 0.05).to(dev)
    outs = []
    for _ in range(2):
        loss, colsum, saved = P.dino_head_forward(xs, xt, vs, one, vt, one, center, ncrops, 10.0, 25.0)
        dx, dv, _ = P.dino_head_backward(saved, torch.ones((), device=dev))
        outs.append((loss.clone(), colsum.clone(), dx.clone(), dv.clone()))
        del saved
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)                      # no atomics, fixed summation orders
    loss, colsum, dx, dv = outs[0]
    rl, rdx, rdv, _, rc1 = O.dino_head_loss_and_grads(xs, xt, vs, one, vt, one, center, ncrops, 0.04, 0.1, round_bf16=True)
    assert abs(float(loss) - float(rl)) <= 1e-3 * abs(float(rl)), (float(loss), float(rl))
    assert maxrel(dx, rdx.cpu()) < 5e-3
    assert maxrel(dv, rdv.cpu()) < 5e-3
    assert_colsum(colsum, xt, vt, one)


def test_dino_head_module_drop_in(P):
    """DINOHead(fused_loss=True) + DINOLoss: the call sequence of lafs_train.py:581-583,600 (head returns what the
    loss consumes), gradients reach the mlp and last_layer.weight_v; same numbers as the unfused modules (reference
    behaviour of DINOHead.forward -> logits -> DINOLoss kernels) and the centre moves identically."""
    torch.manual_seed(5)
    in_dim, K, B, ncrops = 96, 3000, 8, 6
    s_f = P.DINOHead(in_dim, K, nlayers=3, hidden_dim=128, bottleneck_dim=64, fused_loss=True).cuda()
    t_f = P.DINOHead(in_dim, K, nlayers=3, hidden_dim=128, bottleneck_dim=64, fused_loss=True).cuda()
    with torch.no_grad():
        for h in (s_f, t_f):
            h.last_layer.weight_v.normal_(0, 0.5)
    s_u = P.DINOHead(in_dim, K, nlayers=3, hidden_dim=128, bottleneck_dim=64).cuda()
    t_u = P.DINOHead(in_dim, K, nlayers=3, hidden_dim=128, bottleneck_dim=64).cuda()
    s_u.load_state_dict(s_f.state_dict(), strict=True)
    t_u.load_state_dict(t_f.state_dict(), strict=True)
    assert sorted(s_f.state_dict()) == ["last_layer.weight_g", "last_layer.weight_v", "mlp.0.bias", "mlp.0.weight",
                                        "mlp.2.bias", "mlp.2.weight", "mlp.4.bias", "mlp.4.weight"]
    feat_s = torch.randn(ncrops * B, in_dim, device="cuda")
    feat_t = torch.randn(2 * B, in_dim, device="cuda")
    c0 = torch.randn(1, K, device="cuda") * 0.05
    lf = P.DINOLoss(K, ncrops, 0.04, 0.07, 30, 41).cuda()
    lu = P.DINOLoss(K, ncrops, 0.04, 0.07, 30, 41).cuda()
    lf.center = c0.clone()
    lu.center = c0.clone()
    with torch.no_grad():
        to_f, to_u = t_f(feat_t), t_u(feat_t)
    so_f, so_u = s_f(feat_s), s_u(feat_s)
    assert isinstance(so_f, P.DeferredLogits) and so_f.shape == (ncrops * B, K) and torch.is_tensor(so_u)
    loss_f = lf(so_f, to_f, 7)
    loss_u = lu(so_u, to_u, 7)
    loss_f.backward()
    loss_u.backward()
    assert abs(float(loss_f) - float(loss_u)) <= 1e-3 * abs(float(loss_u))
    assert s_f.last_layer.weight_g.grad is None          # norm_last_layer=True: frozen, as in the reference
    for (n, a), (_, b) in zip(s_f.named_parameters(), s_u.named_parameters()):
        if a.grad is None:
            assert b.grad is None, n
            continue
        assert maxrel(a.grad, b.grad) < 2e-2, n          # fp32 logits path vs bf16 tensor-core operands
    torch.testing.assert_close(lf.center, lu.center, rtol=0, atol=2e-3 * float(lu.center.abs().max()) + 1e-4)
    with pytest.raises(TypeError):
        lf(so_f, to_u, 7)
    # no-grad evaluation path: only the teacher's probabilities are ever written
    with torch.no_grad():
        lf.center = c0.clone()
        l2 = lf(s_f(feat_s), t_f(feat_t), 7)
    assert abs(float(l2) - float(loss_f)) <= 1e-6 * abs(float(loss_f))


def test_fused_dino_head_argument_errors(P):
    x = torch.randn(12, 96, device="cuda")
    v = torch.randn(100, 96, device="cuda")
    one = torch.ones(100, device="cuda")
    c = torch.zeros(100, device="cuda")
    with pytest.raises(ValueError):          # bottleneck width must be a multiple of 64
        P.dino_head_forward(x, x[:4], v, one, v, one, c, 6, 10.0, 25.0)
    with pytest.raises(RuntimeError):        # no CPU fallback
        P.dino_head_forward(x.cpu(), x[:4].cpu(), v.cpu(), one.cpu(), v.cpu(), one.cpu(), c.cpu(), 6, 10.0, 25.0)
