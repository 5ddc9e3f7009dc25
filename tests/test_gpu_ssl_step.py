"""GPU: the composed SSL hot-path step (eager) against the oracle, and its CUDA-graph replay
against the eager path (same state evolution over several steps)."""
import copy

import pytest
import torch

from oracle import lafs_oracle as O

pytestmark = pytest.mark.gpu


def build(B=6, L=2, K=2048, dim=128, seed=0):
    import lafs_cvpr2024_b200 as P
    from lafs_cvpr2024_b200.ssl_step import SSLHotPath
    g = torch.Generator().manual_seed(seed)
    host = {
        "img_g": torch.randint(0, 256, (2 * B, 3, 112, 112), generator=g, dtype=torch.uint8),
        "img_l": torch.randint(0, 256, (L * B, 3, 112, 112), generator=g, dtype=torch.uint8),
        "noise_g": torch.randn(2 * B, 196, 2, generator=g) * 5,
        "noise_l": torch.randn(L * B, 196, 2, generator=g) * 5,
        "idx_l": torch.randint(0, 196, (L * B, 36), generator=g),
        "raw_g": torch.randn(2 * B, 392, generator=g),
        "raw_l": torch.randn(L * B, 392, generator=g),
        "student_out": (torch.randn((L + 2) * B, K, generator=g) * 2).bfloat16(),
        "teacher_out": (torch.randn(2 * B, K, generator=g) * 2).bfloat16(),
    }
    shapes = [(1, 197, dim), (dim, 192), (dim,), (33, 7), (5,)]
    sp = [torch.randn(*s, generator=g) * 0.1 for s in shapes]
    tp = [torch.randn(*s, generator=g) * 0.1 for s in shapes]
    return P, SSLHotPath, host, sp, tp


def test_eager_step_matches_oracle():
    P, SSLHotPath, host, sp, tp = build()
    B, L, K = 6, 2, 2048
    dev = {k: v.cuda() for k, v in host.items()}
    spg, tpg = [p.cuda() for p in sp], [p.cuda() for p in tp]
    path = SSLHotPath(K, L, tpg, spg, student_embed=(spg[1], spg[2]), teacher_embed=(tpg[1], tpg[2]))
    center0 = torch.randn(1, K) * 0.1
    path.loss.center = center0.cuda()
    s_g, t_g, s_l = path.landmarks_and_embeddings(dev["raw_g"], dev["noise_g"], dev["img_g"], dev["raw_l"],
                                                  dev["noise_l"], dev["idx_l"], dev["img_l"])
    loss, grad = path.loss_and_grad(dev["student_out"], dev["teacher_out"], 5)
    path.ema_step(0.99)
    # oracle
    img_g = (host["img_g"].float() / 255 - 0.5) / 0.5
    img_l = (host["img_l"].float() / 255 - 0.5) / 0.5
    th_g = O.landmark_post(host["raw_g"], host["noise_g"])
    th_l = O.landmark_post(host["raw_l"], host["noise_l"], host["idx_l"])
    for out, img, th, (w, b) in ((s_g, img_g, th_g, (sp[1], sp[2])), (t_g, img_g, th_g, (tp[1], tp[2])),
                                 (s_l, img_l, th_l, (sp[1], sp[2]))):
        ref = O.gather_embed(img, th, w, b, round_bf16=True)
        assert (out.float().cpu() - ref).abs().max() <= (1e-3 + 2 ** -8) * ref.abs().max()
    temp = float(O.teacher_temp_schedule(0.04, 0.07, 30, 41)[5])
    rl, rg = O.dino_loss_and_grad(host["student_out"].float(), host["teacher_out"].float(), center0, L + 2, temp)
    assert abs(float(loss) - float(rl)) <= 2e-5 * abs(float(rl))
    assert (grad.float().cpu() - rg).abs().max() <= 2 ** -7 * rg.abs().max()
    torch.testing.assert_close(path.loss.center.cpu(), O.dino_center_update(center0, host["teacher_out"].float()),
                               rtol=2e-6, atol=1e-6)
    tp_ref = [p.clone() for p in tp]
    O.ema_update_(tp_ref, sp, 0.99)
    assert all(torch.equal(a, b.cpu()) for a, b in zip(tp_ref, tpg))


def test_graph_replay_equals_eager_over_steps():
    from lafs_cvpr2024_b200.ssl_step import GraphedSSLStep
    P, SSLHotPath, host, sp, tp = build(seed=1)
    B, L, K = 6, 2, 2048
    center0 = torch.randn(1, K) * 0.1

    def fresh():
        spg, tpg = [p.clone().cuda() for p in sp], [p.clone().cuda() for p in tp]
        path = SSLHotPath(K, L, tpg, spg, student_embed=(spg[1], spg[2]), teacher_embed=(tpg[1], tpg[2]))
        path.loss.center = center0.clone().cuda()
        return path, spg, tpg

    dev = {k: v.cuda() for k, v in host.items()}
    # eager: three steps
    path_e, _, tpe = fresh()
    losses_e = []
    for _ in range(3):
        path_e.landmarks_and_embeddings(dev["raw_g"], dev["noise_g"], dev["img_g"], dev["raw_l"], dev["noise_l"],
                                        dev["idx_l"], dev["img_l"])
        loss, grad_e = path_e.loss_and_grad(dev["student_out"], dev["teacher_out"], 5)   # wave-fused, as in the graph
        path_e.ema_step(0.99)
        losses_e.append(float(loss))
    # graph: construction (two warm-up runs + capture) must leave the teacher and the centre untouched
    path_g, _, tpg = fresh()
    g = GraphedSSLStep(path_g, dev, epoch=5, momentum=0.99)
    assert torch.equal(g.center, center0.cuda())
    assert all(torch.equal(a.cpu(), b) for a, b in zip(tpg, tp))
    losses_g = [float(g.replay()) for _ in range(3)]
    assert losses_g == losses_e                      # deterministic kernels: bit-identical
    assert torch.equal(g.out["grad_student"], grad_e)
    assert torch.equal(g.center, path_e.loss.center)
    diffs = [float((a - b).abs().max()) for a, b in zip(tpg, tpe)]
    assert all(d == 0.0 for d in diffs), diffs


def test_student_embed_backward_matches_oracle_autograd():
    """patch_to_embedding weight/bias gradient of the student from the tokens the fused forward kept (uint8 views,
    two view groups accumulated) against fp32 autograd through the oracle's gather + F.linear on bf16-rounded
    tokens and gradients (the operands of the tensor-core GEMM)."""
    P, SSLHotPath, host, sp, tp = build(B=7, L=3, dim=256, seed=2)
    B, L, K, dim = 7, 3, 2048, 256
    dev = {k: v.cuda() for k, v in host.items()}
    spg, tpg = [p.cuda() for p in sp], [p.cuda() for p in tp]
    path = SSLHotPath(K, L, tpg, spg, student_embed=(spg[1], spg[2]), teacher_embed=(tpg[1], tpg[2]))
    path.landmarks_and_embeddings(dev["raw_g"], dev["noise_g"], dev["img_g"], dev["raw_l"], dev["noise_l"], dev["idx_l"],
                                  dev["img_l"], keep_tokens=True)
    g = torch.Generator().manual_seed(9)
    dy_g = (torch.randn(2 * B, 196, dim, generator=g) * 0.05).bfloat16()
    dy_l = (torch.randn(L * B, 36, dim, generator=g) * 0.05).bfloat16()
    for _ in range(2):          # the token buffers are persistent: a second call must give the same result
        gw, gb = path.student_embed_backward(dy_g.cuda(), dy_l.cuda())
    img_g = (host["img_g"].float() / 255 - 0.5) / 0.5
    img_l = (host["img_l"].float() / 255 - 0.5) / 0.5
    th_g = O.landmark_post(host["raw_g"], host["noise_g"])
    th_l = O.landmark_post(host["raw_l"], host["noise_l"], host["idx_l"])
    w = sp[1].clone().requires_grad_(True)
    b = sp[2].clone().requires_grad_(True)
    out = 0.0
    for img, th, dy in ((img_g, th_g, dy_g), (img_l, th_l, dy_l)):
        tok = O.extract_tokens(img, th).bfloat16().float()
        out = out + (torch.nn.functional.linear(tok, w, b) * dy.float()).sum()
    out.backward()
    assert gw.shape == (dim, 192) and gb.shape == (dim,)
    assert (gw.cpu() - w.grad).abs().max() <= 1e-3 * w.grad.abs().max(), float((gw.cpu() - w.grad).abs().max() / w.grad.abs().max())
    assert (gb.cpu() - b.grad).abs().max() <= 1e-3 * b.grad.abs().max()
