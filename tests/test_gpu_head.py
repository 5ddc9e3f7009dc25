"""GPU parity: CosFace / ArcFace margin head (tcgen05 GEMM + fused softmax-CE epilogue)
against golden vectors (reference CosFace) and the CPU oracle."""
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


def bf16_round(x):
    return x.bfloat16().float()


def oracle_on_bf16_operands(x, w, label, kind="cosface", label_b=None, lam=1.0):
    """compare against the oracle evaluated in fp32 on the same bf16-rounded
    normalised operands the tensor cores see."""
    xh = bf16_round(torch.nn.functional.normalize(x))
    wh = bf16_round(torch.nn.functional.normalize(w))
    return O.head_loss_and_grads(xh, wh, label, kind, label_b=label_b, lam=lam, pre_normalized=True)


def make_head(P, cls, w, **kw):
    h = cls(w.shape[1], w.shape[0], None, **kw).cuda()
    with torch.no_grad():
        h.weight.copy_(w)
    return h


def test_cosface_golden_logits_and_loss(P, golden):
    g = golden("cosface")
    x, w, lab = T(g["x"]), T(g["weight"]), T(g["label"])
    h = make_head(P, P.CosFace, w)
    logits = h(x.cuda(), lab.cuda()).cpu()
    ref = T(g["logits_hard"])
    assert logits.shape == ref.shape
    assert (logits - ref).abs().max() <= 1e-3 * ref.abs().max() * 8   # bf16 operands: ~2^-9 * 64
    loss, _ = h.forward_loss_stats(x.cuda(), lab.cuda())
    assert abs(float(loss) - float(g["loss_hard"])) <= 5e-3 * abs(float(g["loss_hard"]))
    # soft (mixup) targets through the dense [B, C] reference API and through the two-label form
    soft = O.mixup_target(lab, w.shape[0], float(g["lam"]))
    logits_s = h(x.cuda(), soft.cuda()).cpu()
    assert (logits_s - T(g["logits_soft"])).abs().max() <= 8e-3 * ref.abs().max()
    loss_s, _ = h.forward_loss_stats(x.cuda(), lab.cuda(), T(g["label_b"]).cuda(), float(g["lam"]))
    assert abs(float(loss_s) - float(g["loss_soft"])) <= 5e-3 * abs(float(g["loss_soft"]))
    loss_d, _ = h.forward_loss_stats(x.cuda(), soft.cuda())
    assert abs(float(loss_d) - float(loss_s)) <= 1e-5 * abs(float(loss_s))


@pytest.mark.parametrize("B,C,D", [(8, 1000, 64), (130, 777, 128), (512, 5000, 512), (64, 3001, 768), (1, 9, 64),
                                   (300, 40000, 512)])
@pytest.mark.parametrize("kind", ["cosface", "arcface"])
def test_head_vs_oracle_on_same_bf16_operands(P, B, C, D, kind):
    torch.manual_seed(B + C + D)
    x = torch.randn(B, D)
    w = torch.randn(C, D) * 0.05
    lab = torch.randint(0, C, (B,))
    lab[0] = C - 1
    cls = P.CosFace if kind == "cosface" else P.ArcFace
    h = make_head(P, cls, w)
    ref_loss, ref_logits, _, _ = oracle_on_bf16_operands(x, w, lab, kind)
    logits = h(x.cuda(), lab.cuda()).cpu()
    scale = ref_logits.abs().max()
    assert (logits - ref_logits).abs().max() <= 1e-3 * scale, float((logits - ref_logits).abs().max())
    loss, lse2 = h.forward_loss_stats(x.cuda(), lab.cuda())
    assert abs(float(loss) - float(ref_loss)) <= 1e-3 * abs(float(ref_loss)), (float(loss), float(ref_loss))
    ref_lse = torch.logsumexp(ref_logits, 1)
    assert (lse2.cpu() * np.log(2.0) - ref_lse).abs().max() <= 1e-3 * ref_lse.abs().max()


def test_head_mixup_soft_labels_vs_oracle(P):
    torch.manual_seed(5)
    B, C, D = 257, 3333, 256
    x, w = torch.randn(B, D), torch.randn(C, D)
    lab = torch.randint(0, C, (B,))
    lam = 0.37
    h = make_head(P, P.CosFace, w)
    ref_loss, ref_logits, _, _ = oracle_on_bf16_operands(x, w, lab, "cosface", label_b=lab.flip(0), lam=lam)
    loss, _ = h.forward_loss_stats(x.cuda(), lab.cuda(), lab.flip(0).cuda(), lam)
    assert abs(float(loss) - float(ref_loss)) <= 1e-3 * abs(float(ref_loss))
    dense = O.mixup_target(lab, C, lam)
    logits = h(x.cuda(), dense.cuda()).cpu()
    assert (logits - ref_logits).abs().max() <= 1e-3 * ref_logits.abs().max()


def test_shard_mapping_bit_exact(P, golden):
    for row in golden("shards")["table"]:
        C, R, sizes = int(row[0]), int(row[1]), [int(v) for v in row[2:] if v > 0]
        b = [hi - lo for lo, hi in P.shard_bounds(C, R) if hi > lo]
        assert b == sizes
    lab = torch.tensor([0, 11678, 11679, 93430])
    sh, loc = P.label_to_shard(lab, 93431, 8)
    assert sh.tolist() == [0, 0, 1, 7] and loc.tolist() == [0, 11678, 0, 93430 - 7 * 11679]


def test_sharded_stats_merge_equals_unsharded(P):
    """Class-parallel path on one GPU: every shard's partial statistics merged == full softmax."""
    torch.manual_seed(7)
    B, C, D, R = 96, 10007, 512, 8
    x, w = torch.randn(B, D), torch.randn(C, D)
    lab = torch.randint(0, C, (B,))
    full = make_head(P, P.CosFace, w)
    loss_full, lse_full = full.forward_loss_stats(x.cuda(), lab.cuda())
    from lafs_cvpr2024_b200 import _lib
    parts = []
    for r, (lo, hi) in enumerate(P.shard_bounds(C, R)):
        h = P.CosFace(D, C, None, shard=(r, R)).cuda()
        assert h.weight.shape[0] == hi - lo
        with torch.no_grad():
            h.weight.copy_(w[lo:hi])
        st, _ = h.forward_stats(x.cuda(), lab.cuda())
        parts.append(st)
    parts = torch.stack(parts).contiguous()
    merged = torch.empty(B, 4, device="cuda")
    _lib.call("lafs_head_merge", parts.data_ptr(), R, B, merged.data_ptr(), _lib.stream())
    loss = torch.empty((), device="cuda"); lse2 = torch.empty(B, device="cuda")
    la = lab.cuda()
    _lib.call("lafs_head_loss", merged.data_ptr(), la.data_ptr(), None, 1.0, B, lse2.data_ptr(), loss.data_ptr(), _lib.stream())
    assert abs(float(loss) - float(loss_full)) <= 1e-5 * abs(float(loss_full))
    torch.testing.assert_close(lse2, lse_full, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,C,D", [(8, 1000, 64), (130, 777, 128), (512, 5000, 512), (64, 3001, 768), (300, 20011, 512)])
@pytest.mark.parametrize("kind", ["cosface", "arcface"])
def test_head_backward_vs_oracle(P, B, C, D, kind):
    """dE, dW of the fused loss against autograd on the oracle (fp32, same bf16-rounded operands)."""
    torch.manual_seed(B + C + D + 1)
    x = torch.randn(B, D)
    w = torch.randn(C, D) * 0.05
    lab = torch.randint(0, C, (B,))
    cls = P.CosFace if kind == "cosface" else P.ArcFace
    h = make_head(P, cls, w)
    ref_loss, _, ref_gx, ref_gw = O.head_loss_and_grads(x, w, lab, kind)
    xg = x.cuda().requires_grad_(True)
    loss = h.forward_loss(xg, lab.cuda())
    loss.backward()
    assert abs(float(loss) - float(ref_loss)) <= 5e-3 * abs(float(ref_loss))
    gx, gw = xg.grad.cpu(), h.weight.grad.cpu()
    # bf16 operands + bf16 logit gradient: compare in the max-norm-relative sense
    assert (gx - ref_gx).abs().max() <= 2e-2 * ref_gx.abs().max(), float((gx - ref_gx).abs().max() / ref_gx.abs().max())
    assert (gw - ref_gw).abs().max() <= 2e-2 * ref_gw.abs().max(), float((gw - ref_gw).abs().max() / ref_gw.abs().max())
    # direction check: cosine similarity of the full gradients
    cs = torch.nn.functional.cosine_similarity(gw.flatten(), ref_gw.flatten(), dim=0)
    assert cs > 0.9995, float(cs)


def test_head_backward_soft_labels_and_grad_out(P):
    torch.manual_seed(11)
    B, C, D = 96, 2048, 256
    x, w = torch.randn(B, D), torch.randn(C, D) * 0.1
    lab = torch.randint(0, C, (B,))
    lam = 0.3
    h = make_head(P, P.CosFace, w)
    ref_loss, _, ref_gx, ref_gw = O.head_loss_and_grads(x, w, lab, "cosface", label_b=lab.flip(0), lam=lam, grad_out=3.0)
    xg = x.cuda().requires_grad_(True)
    loss = h.forward_loss(xg, lab.cuda(), lab.flip(0).cuda(), lam)
    (loss * 3.0).backward()
    gx, gw = xg.grad.cpu(), h.weight.grad.cpu()
    assert (gx - ref_gx).abs().max() <= 2e-2 * ref_gx.abs().max()
    assert (gw - ref_gw).abs().max() <= 2e-2 * ref_gw.abs().max()


def test_cosface_forward_logits_backward_golden(P, golden):
    """Drop-in use: logits = head(x, label); CrossEntropyLoss(logits, label).backward()  (reference grads)."""
    g = golden("cosface")
    x, w, lab = T(g["x"]), T(g["weight"]), T(g["label"])
    h = make_head(P, P.CosFace, w)
    xg = x.cuda().requires_grad_(True)
    logits = h(xg, lab.cuda())
    torch.nn.CrossEntropyLoss()(logits, lab.cuda()).backward()
    ref_gx, ref_gw = T(g["grad_x_hard"]), T(g["grad_w_hard"])
    assert (xg.grad.cpu() - ref_gx).abs().max() <= 2e-2 * ref_gx.abs().max()
    assert (h.weight.grad.cpu() - ref_gw).abs().max() <= 2e-2 * ref_gw.abs().max()


@pytest.mark.parametrize("shard", [None, (1, 3)])
def test_arcface_forward_logits_backward_vs_oracle(P, shard):
    """logits = ArcFace(x, label); CrossEntropyLoss(logits, label).backward(): the target column's d phi / d cos factor
    of the full-logits path (computed on the device without a host sync), unsharded and on a class shard (some
    labels owned by other ranks), against autograd through the oracle on the same bf16 operands."""
    torch.manual_seed(77)
    B, C, D = 96, 1501, 128
    x, w = torch.randn(B, D), torch.randn(C, D) * 0.05
    lab = torch.randint(0, C, (B,))
    lab[0], lab[1] = 0, C - 1
    lo, hi = (0, C) if shard is None else P.shard_bounds(C, shard[1])[shard[0]]
    h = P.ArcFace(D, C, None, shard=shard).cuda()
    with torch.no_grad():
        h.weight.copy_(w[lo:hi])
    xg = x.cuda().requires_grad_(True)
    logits = h(xg, lab.cuda())
    up = torch.randn(B, hi - lo, generator=torch.Generator().manual_seed(5))
    (logits * up.cuda()).sum().backward()
    xh = bf16_round(torch.nn.functional.normalize(x)).requires_grad_(True)
    wh = bf16_round(torch.nn.functional.normalize(w)).requires_grad_(True)
    ref = O.arcface_logits(xh, wh, lab, pre_normalized=True)[:, lo:hi]
    assert (logits.detach().cpu() - ref.detach()).abs().max() <= 1e-3 * ref.detach().abs().max()
    (ref * up).sum().backward()
    # chain through F.normalize on both sides (the oracle differentiated w.r.t. the unit operands)
    xn, wn = x.norm(dim=1, keepdim=True), w.norm(dim=1, keepdim=True)
    xu, wu = x / xn, w / wn
    ref_gx = (xh.grad - xu * (xu * xh.grad).sum(1, keepdim=True)) / xn
    ref_gw = ((wh.grad - wu * (wu * wh.grad).sum(1, keepdim=True)) / wn)[lo:hi]
    assert (xg.grad.cpu() - ref_gx).abs().max() <= 1e-2 * ref_gx.abs().max()
    assert (h.weight.grad.cpu() - ref_gw).abs().max() <= 1e-2 * ref_gw.abs().max()


def test_label_range_check_is_opt_in(P):
    h = P.CosFace(64, 100, None).cuda()
    x = torch.randn(4, 64, device="cuda")
    bad = torch.tensor([1, 2, 100, 3], device="cuda")
    assert torch.isfinite(h.forward_loss(x, bad))            # default: the out-of-range row contributes its lse only
    h.check_labels = True
    with pytest.raises(IndexError):
        h.forward_loss(x, bad)
    assert torch.isfinite(h.forward_loss(x, torch.tensor([1, 2, 99, 0], device="cuda")))


def _step_outputs(P, B, C, D, kind, env):
    """loss, row lse, dE, dW of one fused-loss step under the given kernel-variant switches."""
    import os
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        torch.manual_seed(1234)
        x = torch.randn(B, D)
        w = torch.randn(C, D) * 0.05
        lab = torch.randint(0, C, (B,))
        cls = P.CosFace if kind == "cosface" else P.ArcFace
        h = make_head(P, cls, w)
        xg = x.cuda().requires_grad_(True)
        loss = h.forward_loss(xg, lab.cuda())
        loss.backward()
        logits = h(x.cuda(), lab.cuda())
        torch.cuda.synchronize()
        return float(loss), xg.grad.cpu(), h.weight.grad.cpu(), logits.cpu()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("B,C,D", [(512, 9001, 512), (256, 4099, 256), (130, 3000, 128), (1024, 2500, 512)])
@pytest.mark.parametrize("kind", ["cosface", "arcface"])
def test_head_kernel_variants_agree(P, B, C, D, kind):
    """CTA-pair (cta_group::2) forward/grad GEMMs vs the single-CTA kernel, dE with / without W_hat multicast: same
    math, so the results agree to fp32 accumulation-order noise."""
    import os
    os.environ["LAFS_DW_DIAG"] = "0"           # the fused / unfused switches select among the non-default dW kernels
    try:
        _variants_agree(P, B, C, D, kind)
    finally:
        os.environ.pop("LAFS_DW_DIAG", None)


def _variants_agree(P, B, C, D, kind):
    base = _step_outputs(P, B, C, D, kind, {"LAFS_HEAD_1SM": "1"})
    for env in ({"LAFS_HEAD_1SM": "0"},
                {"LAFS_HEAD_1SM": "0", "LAFS_DE_CLUSTER": "1"}, {"LAFS_HEAD_1SM": "0", "LAFS_DE_CLUSTER": "2"}):
        out = _step_outputs(P, B, C, D, kind, env)
        assert abs(out[0] - base[0]) <= 1e-5 * abs(base[0]), (env, out[0], base[0])
        assert (out[3] - base[3]).abs().max() <= 1e-4, (env, float((out[3] - base[3]).abs().max()))
        for a, b in ((out[1], base[1]), (out[2], base[2])):
            assert (a - b).abs().max() <= 2e-3 * b.abs().max() + 1e-9, (env, float((a - b).abs().max() / b.abs().max()))


@pytest.mark.parametrize("B,C,D", [(512, 9001, 512), (130, 3000, 128), (300, 4099, 256), (64, 1000, 768)])
@pytest.mark.parametrize("kind", ["cosface", "arcface"])
def test_dw_jacobian_on_tensor_core_variant(P, B, C, D, kind):
    """Default dW path (per-class dots from the gradient kernel + dW = inv * (G^T.E - diag(t).W_hat) in one GEMM)
    against GEMM + normalize_bwd pass, LAFS_DW_DIAG=0 (t is rounded to bf16: ~6e-4 max-norm relative on dW)."""
    base = _step_outputs(P, B, C, D, kind, {"LAFS_DW_DIAG": "0"})
    out = _step_outputs(P, B, C, D, kind, {"LAFS_DW_DIAG": "1"})
    assert abs(out[0] - base[0]) <= 1e-6 * abs(base[0])
    assert (out[1] - base[1]).abs().max() <= 1e-5 * base[1].abs().max()            # dE: untouched path
    assert (out[2] - base[2]).abs().max() <= 5e-3 * base[2].abs().max(), float((out[2] - base[2]).abs().max() / base[2].abs().max())


# ---- the exact shapes bench.py times (BASELINE configs[2], configs[3]) and their 8-way shard-local sizes ----
def _bench_case(P, kind, B, C, D, shard=None, seed=3):
    """One fused-loss step at a bench shape against the oracle on the SAME bf16-rounded unit operands
    (fp32 math, gradients through the normalisation Jacobians).  shard=(rank, world): only that class
    slice lives on the GPU (class_lo != 0), the oracle runs the full problem."""
    torch.manual_seed(seed)
    x = torch.randn(B, D)
    w = torch.empty(C, D)
    torch.nn.init.xavier_uniform_(w)                      # the reference's init (ViT_face.py:46-47)
    lab = torch.randint(0, C, (B,))
    lab[0], lab[1] = C - 1, 0
    cls = P.CosFace if kind == "cosface" else P.ArcFace
    ref_loss, ref_logits, ref_gx, ref_gw = O.head_loss_and_grads(x, w, lab, kind, pre_normalized="bf16_st")
    ref_lse = torch.logsumexp(ref_logits, 1)
    del ref_logits
    if shard is None:
        h = make_head(P, cls, w)
        xg = x.cuda().requires_grad_(True)
        loss = h.forward_loss(xg, lab.cuda())
        loss.backward()
        assert abs(float(loss) - float(ref_loss)) <= 1e-3 * abs(float(ref_loss)), (float(loss), float(ref_loss))
        gx, gw = xg.grad.cpu(), h.weight.grad.cpu()
        lo, hi = 0, C
    else:
        r, R = shard
        lo, hi = P.shard_bounds(C, R)[r]
        h = cls(D, C, None, shard=(r, R)).cuda()
        assert h.class_lo == lo and h.weight.shape[0] == hi - lo and lo != 0
        with torch.no_grad():
            h.weight.copy_(w[lo:hi])
        # per-row statistics of this shard against the oracle's logits restricted to the shard
        stats, (la, lb, lam, e_hat, w_hat) = h.forward_stats(x.cuda(), lab.cuda())
        xs = torch.nn.functional.normalize(x).bfloat16().float()
        wsl = torch.nn.functional.normalize(w[lo:hi]).bfloat16().float()
        z = 64.0 * (xs @ wsl.t())
        own = (lab >= lo) & (lab < hi)
        logit_fn = O.cosface_logits if kind == "cosface" else O.arcface_logits
        z[own] = logit_fn(xs[own], wsl, lab[own] - lo, pre_normalized=True)     # margin on the rows this shard owns
        st = stats.cpu()
        lse_shard = st[:, 0] * np.log(2.0) + torch.log(st[:, 1])
        assert (lse_shard - torch.logsumexp(z, 1)).abs().max() <= 1e-3
        assert (st[own, 2] - z[own, (lab[own] - lo)]).abs().max() <= 2e-3 * 64
        # gradient of this shard given the FULL row lse (what the merged statistics deliver)
        from lafs_cvpr2024_b200 import _lib, margin_head as MH
        lse2 = (ref_lse / np.log(2.0)).float().cuda().contiguous()
        ldg = MH._round8(hi - lo)
        G = torch.empty(B, ldg, dtype=torch.bfloat16, device="cuda")
        one = torch.ones((), device="cuda")
        _lib.call("lafs_head_grad_logits", e_hat.data_ptr(), w_hat.data_ptr(), la.data_ptr(), None, 1.0, B, hi - lo, D, lo,
                  float(h.s), float(h.m), h.kind, lse2.data_ptr(), one.data_ptr(), float(h.s) / B, G.data_ptr(), ldg,
                  _lib.stream())
        _, inv_e = MH._prep(x.cuda(), want_inv=True)
        _, inv_w = MH._prep(h.weight, want_inv=True)
        de_part, gw = MH._head_backward(G, ldg, e_hat, w_hat, inv_e, inv_w, B, hi - lo, D, False)
        gw = gw.cpu()
        gx = None      # dE of one shard is a partial sum: checked through dW and the statistics here
    sl = slice(lo, hi)
    if gx is not None:
        err = float((gx - ref_gx).abs().max() / ref_gx.abs().max())
        assert err <= 5e-3, ("dE", err)
    # dW: the rows of the target classes (largest entries) plus a sampled 4k-class slice
    idx = torch.unique(torch.cat([lab[(lab >= lo) & (lab < hi)], torch.randint(lo, hi, (4096,))]))
    err = float((gw[idx - lo] - ref_gw[idx]).abs().max() / ref_gw[sl].abs().max())
    assert err <= 5e-3, ("dW", err)
    cs = torch.nn.functional.cosine_similarity(gw.flatten(), ref_gw[sl].flatten(), dim=0)
    assert cs > 0.9999, float(cs)


@pytest.mark.parametrize("kind,B,C,D", [("cosface", 512, 93431, 512), ("arcface", 1024, 205990, 512),
                                        ("cosface", 512, 93431, 768)])
def test_head_bench_shapes_vs_oracle(P, kind, B, C, D):
    _bench_case(P, kind, B, C, D)


@pytest.mark.parametrize("kind,B,C,rank", [("cosface", 512, 93431, 3), ("cosface", 512, 93431, 7),
                                           ("arcface", 1024, 205990, 5), ("arcface", 1024, 205990, 7)])
def test_head_bench_shard_local_shapes_vs_oracle(P, kind, B, C, rank):
    """8-way class shards of the bench shapes: 11,679 / 11,678 and 25,749 / 25,747 local classes, class_lo != 0."""
    _bench_case(P, kind, B, C, 512, shard=(rank, 8))
