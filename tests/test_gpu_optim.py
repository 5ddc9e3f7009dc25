"""GPU parity: the fused student update (per-tensor clip + AdamW + teacher EMA, csrc/optim.cu) against the oracle
(utils.clip_gradients restated + torch.optim.AdamW + the EMA loop, on the host in fp32).  Tolerance 1e-5 relative
on the parameters (two fp32 evaluations of the same formulas; FMA contraction on the host).  The per-tensor norms
are compared at 5e-5: torch's fp32 CPU reduction is itself 2.1e-5 off the float64 value on a 2112x768 tensor (the
kernel's fixed-order chunked sum is the closer one), and that relative error carries into the clip coefficient and
the moments of clipped tensors.  The north star's bound for updated / EMA'd weights is 1e-3."""
import numpy as np
import pytest
import torch

from oracle import lafs_oracle as O

pytestmark = pytest.mark.gpu


def _case(shapes, reg, steps, clip, seed=0, cancel=None, teacher=True):
    import lafs_cvpr2024_b200 as P
    g = torch.Generator().manual_seed(seed)
    p0 = [torch.randn(*s, generator=g) * 0.05 for s in shapes]
    k0 = [torch.randn(*s, generator=g) * 0.05 for s in shapes]
    # oracle side
    p_ref = [torch.nn.Parameter(p.clone()) for p in p0]
    k_ref = [k.clone() for k in k0] if teacher else None
    groups = [{"params": [p for p, r in zip(p_ref, reg) if r]}, {"params": [p for p, r in zip(p_ref, reg) if not r], "weight_decay": 0.}]
    opt = torch.optim.AdamW(groups)
    # kernel side
    p_gpu = [p.clone().cuda() for p in p0]
    k_gpu = [k.clone().cuda() for k in k0] if teacher else None
    upd = P.StudentUpdate(p_gpu, k_gpu, regularized=reg)
    lr_s = O.cosine_scheduler(5e-4, 1e-6, 1, steps)
    wd_s = O.cosine_scheduler(0.04, 0.4, 1, steps)
    mo_s = O.cosine_scheduler(0.996, 1.0, 1, steps)
    for it in range(steps):
        grads = [torch.randn(*s, generator=g) * (3.0 if i % 3 == 0 else 0.02) for i, s in enumerate(shapes)]
        if cancel is not None:
            for i in cancel(it):
                grads[i] = None
        norms_ref = O.student_update_(p_ref, grads, k_ref, opt, float(lr_s[it]), float(wd_s[it]), clip, mo_s[it])
        norms = upd.step([None if x is None else x.cuda() for x in grads], float(lr_s[it]), float(wd_s[it]), clip,
                         mo_s[it] if teacher else None)
        if clip:
            have = [i for i, x in enumerate(grads) if x is not None]
            got = norms.cpu()[have]
            np.testing.assert_allclose(got.numpy(), np.array(norms_ref, dtype=np.float32), rtol=5e-5)
    for i, (a, b) in enumerate(zip(p_ref, p_gpu)):
        torch.testing.assert_close(b.cpu(), a.detach(), rtol=1e-5, atol=1e-8, msg=lambda m, i=i: f"param {i}: {m}")
    for i, p in enumerate(p_ref):
        st = opt.state.get(p, None)
        if st:
            # exp_avg is a signed running mean: an element that cancels to ~0 carries the absolute rounding error of
            # its terms, so the absolute tolerance scales with the tensor (2e-6 of its largest entry)
            torch.testing.assert_close(upd.exp_avg[i].cpu(), st["exp_avg"], rtol=5e-5, atol=2e-6 * float(st["exp_avg"].abs().max()))
            torch.testing.assert_close(upd.exp_avg_sq[i].cpu(), st["exp_avg_sq"], rtol=1e-4,
                                       atol=1e-7 * float(st["exp_avg_sq"].abs().max()))
    if teacher:
        for a, b in zip(k_ref, k_gpu):
            torch.testing.assert_close(b.cpu(), a, rtol=1e-5, atol=1e-8)
    return upd


SHAPES = [(768, 192), (768,), (1, 197, 768), (2112, 768), (33, 7), (5,), (16385,), (3000, 256), (1,)]
REG = [True, False, True, True, True, False, False, True, False]


def test_student_update_matches_oracle_over_steps():
    _case(SHAPES, REG, steps=4, clip=3.0)


def test_student_update_without_clip_and_without_teacher():
    _case(SHAPES, REG, steps=2, clip=0.0)
    _case(SHAPES[:4], REG[:4], steps=2, clip=0.3, teacher=False)


def test_cancelled_gradients_are_skipped_like_adamw_skips_none():
    """utils.cancel_gradients_last_layer sets p.grad = None during the first epoch: AdamW then skips the tensor
    (no decay, no moment update) while the teacher EMA still runs; a tensor that never gets a gradient (the unused
    CosFace weight of the SSL student) is only EMA'd."""
    upd = _case(SHAPES, REG, steps=3, clip=3.0, cancel=lambda it: [3, 7])
    assert upd.steps[3] == 0 and upd.steps[0] == 3
    assert float(upd.exp_avg[3].abs().max()) == 0.0


def test_unfreezing_later_needs_a_split_call():
    import lafs_cvpr2024_b200 as P
    p = [torch.randn(10, 10).cuda(), torch.randn(7).cuda()]
    upd = P.StudentUpdate(p, None, regularized=[True, False])
    upd.step([torch.randn(10, 10).cuda(), None], 1e-3, 0.04, 3.0)
    with pytest.raises(ValueError):
        upd.step([torch.randn(10, 10).cuda(), torch.randn(7).cuda()], 1e-3, 0.04, 3.0)


def test_student_update_full_vit_b_list_is_deterministic_and_finite():
    """The list bench.py times (147 tensors, 110 M parameters): two identical runs give identical bits, norms match
    torch.linalg.vector_norm per tensor."""
    import bench
    import lafs_cvpr2024_b200 as P
    shapes = bench.vit_param_shapes("B")
    g = torch.Generator(device="cuda").manual_seed(1)
    base = [torch.randn(*s, device="cuda", generator=g) * 0.02 for s in shapes]
    grads = [torch.randn(*s, device="cuda", generator=g) * 0.01 for s in shapes]
    outs = []
    for _ in range(2):
        p = [b.clone() for b in base]
        k = [b.clone() for b in base]
        upd = P.StudentUpdate(p, k, regularized=[len(s) > 1 for s in shapes])
        norms = upd.step(grads, 5e-4, 0.04, 3.0, 0.996).clone()
        outs.append((p, k, norms))
    for a, b in zip(outs[0][0] + outs[0][1], outs[1][0] + outs[1][1]):
        assert torch.equal(a, b) and bool(torch.isfinite(a).all())
    ref = torch.stack([torch.linalg.vector_norm(x) for x in grads])
    torch.testing.assert_close(outs[0][2], ref, rtol=1e-5, atol=0)
