"""GPU parity: teacher EMA and DINO loss through the package -> C ABI -> sm_100a kernels,
against the committed golden vectors (reference outputs) and the CPU oracle."""
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


# ------------------------------------------------------------------------------- EMA
def test_ema_golden_bit_exact(P, golden):
    g = golden("ema")
    n = sum(1 for k in g if k.startswith("q"))
    q = [T(g[f"q{i}"]).cuda() for i in range(n)]
    k = [T(g[f"k0_{i}"]).cuda() for i in range(n)]
    P.ema_update_(k, q, float(g["m"]))
    for i in range(n):
        assert torch.equal(k[i].cpu(), T(g[f"k1_{i}"])), i


def test_ema_vit_b_shapes_vs_oracle(P):
    # the 15 distinct parameter shapes of the ViT-B student + DINO head, a few each,
    # plus odd sizes and a deliberately 4-byte-aligned (not 16-byte) view
    torch.manual_seed(0)
    shapes = [(256,), (768,), (2048,), (2112, 768), (768, 704), (2048, 768), (768, 2048), (1, 197, 768),
              (1, 1, 768), (768, 192), (2048, 2048), (256, 2048), (65536, 1), (1000, 256), (3000, 768),
              (1,), (5,), (16385,), (33, 77)]
    q = [torch.randn(*s) for s in shapes]
    k = [torch.randn(*s) for s in shapes]
    base_k, base_q = torch.randn(4099), torch.randn(4099)
    qg = [a.cuda() for a in q] + [base_q.cuda()[1:]]
    kg = [a.cuda() for a in k] + [base_k.cuda()[1:]]
    q.append(base_q[1:]); k.append(base_k[1:].clone())
    sched = O.cosine_scheduler(0.996, 1, 41, 1000)
    for it in (0, 12345, len(sched) - 1):
        m = sched[it]
        O.ema_update_(k, q, m)
        P.ema_update_(kg, qg, m)
    for a, b in zip(k, kg):
        assert torch.equal(a, b.cpu())
    P.ema_update_([], [], 0.5)  # empty list is a no-op


def test_ema_full_vit_b_parameter_list_bit_exact(P):
    """The list bench.py times: 147 tensors / 110,261,248 fp32 parameters (ViT-B backbone incl. the unused
    CosFace weight + DINOHead), one launch, bit-exact against lafs_train.py:610-613 on the host."""
    import bench
    shapes = bench.vit_param_shapes("B")
    assert len(shapes) == 147 and sum(int(np.prod(s)) for s in shapes) == 110261248
    g = torch.Generator().manual_seed(1)
    q = [torch.randn(*s, generator=g) * 0.02 for s in shapes]
    k = [torch.randn(*s, generator=g) * 0.02 for s in shapes]
    qg, kg = [a.cuda() for a in q], [a.cuda() for a in k]
    m = O.cosine_scheduler(0.996, 1, 41, 1000)[777]
    O.ema_update_(k, q, m)
    P.ema_update_(kg, qg, m)
    for i, (a, b) in enumerate(zip(k, kg)):
        assert torch.equal(a, b.cpu()), (i, shapes[i])


def test_ema_idempotent_at_m1_and_copies_at_m0(P):
    k, q = torch.randn(100003, device="cuda"), torch.randn(100003, device="cuda")
    k0 = k.clone()
    P.ema_update_([k], [q], 1.0)
    assert torch.equal(k, k0)
    P.ema_update_([k], [q], 0.0)
    assert torch.equal(k, q)


# ------------------------------------------------------------------------------- DINO
def run_dino(P, s, t, center, ncrops, epoch, sched=(0.04, 0.07, 30, 41)):
    crit = P.DINOLoss(s.shape[1], ncrops, *sched).cuda()
    crit.center = center.clone().cuda()
    sg = s.cuda().requires_grad_(True)
    loss = crit(sg, t.cuda(), epoch)
    loss.backward()
    return loss.detach().cpu(), sg.grad.detach().cpu(), crit.center.detach().cpu()


def test_dino_golden_fp32(P, golden):
    g = golden("dino")
    loss, grad, c1 = run_dino(P, T(g["student"]), T(g["teacher"]), T(g["center0"]), int(g["ncrops"]), int(g["epoch"]))
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    gref = T(g["grad_student"])
    assert (grad - gref).abs().max() <= 1e-3 * gref.abs().max()          # north-star tolerance
    assert (grad - gref).abs().max() <= 2e-5 * gref.abs().max()          # what fp32 math achieves
    torch.testing.assert_close(c1, T(g["center1"]), rtol=1e-6, atol=1e-7)


def test_dino_golden_bf16_inputs(P, golden):
    g = golden("dino_bf16")
    s = T(g["student"]).view(torch.bfloat16)
    t = T(g["teacher"]).view(torch.bfloat16)
    loss, grad, c1 = run_dino(P, s, t, T(g["center0"]), int(g["ncrops"]), int(g["epoch"]))
    assert abs(float(loss) - float(g["loss"])) <= 1e-5 * abs(float(g["loss"]))
    gref = T(g["grad_student"])
    assert grad.dtype == torch.bfloat16
    # gradient is rounded to bf16 on store: 2^-9 relative per element
    assert (grad.float() - gref).abs().max() <= 1e-3 * gref.abs().max() + 2 ** -8 * gref.abs().max()
    torch.testing.assert_close(c1, T(g["center1"]), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("B,K,ncrops", [(3, 8, 2), (5, 1000, 3), (8, 4096, 6), (16, 65536, 10), (2, 100000, 12),
                                        (33, 520, 7)])
def test_dino_vs_oracle(P, dtype, B, K, ncrops):
    torch.manual_seed(B * 1000 + K + ncrops)
    s = (torch.randn(ncrops * B, K) * 3).to(dtype)
    t = (torch.randn(2 * B, K) * 3).to(dtype)
    center = torch.randn(1, K) * 0.5
    epoch = 7
    temp = float(O.teacher_temp_schedule(0.04, 0.07, 30, 41)[epoch])
    ref_loss, ref_grad = O.dino_loss_and_grad(s.float(), t.float(), center, ncrops, temp)
    ref_c = O.dino_center_update(center, t.float())
    loss, grad, c1 = run_dino(P, s, t, center, ncrops, epoch)
    assert abs(float(loss) - float(ref_loss)) <= 2e-5 * abs(float(ref_loss)), (float(loss), float(ref_loss))
    tol = 2e-5 if dtype == torch.float32 else 2 ** -7
    assert (grad.float() - ref_grad).abs().max() <= tol * ref_grad.abs().max()
    torch.testing.assert_close(c1, ref_c, rtol=2e-6, atol=1e-6)


def test_dino_extreme_logits_are_stable(P):
    # max-subtraction must hold for huge logits and a far-away centre
    torch.manual_seed(1)
    B, K, ncrops = 4, 2048, 4
    s = torch.randn(ncrops * B, K) * 50 + 300
    t = torch.randn(2 * B, K) * 20 - 500
    center = torch.full((1, K), -480.0)
    temp = float(O.teacher_temp_schedule(0.04, 0.07, 30, 41)[0])
    ref_loss, ref_grad = O.dino_loss_and_grad(s, t, center, ncrops, temp)
    loss, grad, _ = run_dino(P, s, t, center, ncrops, 0)
    assert torch.isfinite(loss)
    assert abs(float(loss) - float(ref_loss)) <= 1e-4 * abs(float(ref_loss))
    assert (grad - ref_grad).abs().max() <= 1e-3 * ref_grad.abs().max()


def test_dino_full_size_properties(P):
    """BASELINE config 2 shape (B=256, K=65536, ncrops=6, bf16): size-independent properties."""
    torch.manual_seed(2)
    B, K, ncrops = 256, 65536, 6
    s = torch.randn(ncrops * B, K, device="cuda", dtype=torch.bfloat16)
    t = torch.randn(2 * B, K, device="cuda", dtype=torch.bfloat16)
    crit = P.DINOLoss(K, ncrops, 0.04, 0.07, 30, 41).cuda()
    crit.center = torch.randn(1, K, device="cuda") * 0.1
    c0 = crit.center.clone()
    sg = s.clone().requires_grad_(True)
    loss = crit(sg, t, 10)
    loss.backward()
    # (a) every gradient row sums to zero:  n_v*sum(p) - sum_{iq != v} sum(q) = 0
    #     (up to the bf16 rounding of the stored gradient: 2^-9 relative per element)
    rs = sg.grad.float().sum(1)
    assert rs.abs().max() < 2 ** -8 * sg.grad.float().abs().sum(1).max()
    # (b) centre update equals the closed form from an independent column sum
    ref_c = c0 * 0.9 + (t.float().sum(0, keepdim=True) / (2 * B)) * 0.1
    torch.testing.assert_close(crit.center, ref_c, rtol=1e-4, atol=1e-6)
    # (c) loss of a subset computed by the oracle on CPU matches the same rows on the GPU
    idx = torch.arange(0, B, 32)
    rows_s = torch.cat([s[v * B + idx] for v in range(ncrops)]).cpu()
    rows_t = torch.cat([t[iq * B + idx] for iq in range(2)]).cpu()
    temp = float(crit.teacher_temp_schedule[10])
    ref = O.dino_loss(rows_s.float(), rows_t.float(), c0.cpu(), ncrops, temp)
    crit2 = P.DINOLoss(K, ncrops, 0.04, 0.07, 30, 41).cuda()
    crit2.center = c0.clone()
    sub = crit2(rows_s.cuda(), rows_t.cuda(), 10)
    assert abs(float(sub) - float(ref)) <= 2e-5 * abs(float(ref))
    # (d) determinism: same inputs, same bits
    crit3 = P.DINOLoss(K, ncrops, 0.04, 0.07, 30, 41).cuda()
    crit3.center = c0.clone()
    assert float(crit3(s, t, 10)) == float(loss)


def test_update_center_standalone(P):
    torch.manual_seed(3)
    K = 4096
    crit = P.DINOLoss(K, 4, 0.04, 0.07, 30, 41).cuda()
    t = torch.randn(10, K, device="cuda")
    crit.update_center(t)
    ref = O.dino_center_update(torch.zeros(1, K), t.cpu())
    torch.testing.assert_close(crit.center.cpu(), ref, rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("B,K,ncrops,dtype", [(7, 1000, 3, torch.float32), (96, 4096, 6, torch.bfloat16),
                                               (256, 65536, 6, torch.bfloat16)])
def test_dino_wave_fused_fwd_bwd_equals_separate(P, B, K, ncrops, dtype):
    """loss_and_grad (one fused call) == forward() + backward()."""
    torch.manual_seed(B + K)
    s = (torch.randn(ncrops * B, K, device="cuda") * 2).to(dtype)
    t = (torch.randn(2 * B, K, device="cuda") * 2).to(dtype)
    c0 = torch.randn(1, K, device="cuda") * 0.3
    a = P.DINOLoss(K, ncrops, 0.04, 0.07, 30, 41).cuda(); a.center = c0.clone()
    b = P.DINOLoss(K, ncrops, 0.04, 0.07, 30, 41).cuda(); b.center = c0.clone()
    sg = s.clone().requires_grad_(True)
    la = a(sg, t, 4)
    la.backward()
    lb, gb = b.loss_and_grad(s, t, 4)
    assert float(la) == float(lb)
    assert torch.equal(sg.grad, gb)
    torch.testing.assert_close(a.center, b.center, rtol=1e-5, atol=1e-6)   # column sums merge in a different order
    sc = torch.full((), 3.0, device="cuda")
    b.center = c0.clone()
    lc, gc = b.loss_and_grad(s, t, 4, grad_scale=sc)
    torch.testing.assert_close(gc.float(), gb.float() * 3.0, rtol=2 ** -7, atol=1e-30)   # bf16 denormals
