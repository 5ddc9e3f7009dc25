"""CPU: the oracle restatement against the committed golden vectors that the reference
itself produced (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import torch

from oracle import lafs_oracle as O


def T(a):
    return torch.from_numpy(np.asarray(a))


def close(a, b, rtol=2e-5, atol=1e-7):
    torch.testing.assert_close(a, b, rtol=rtol, atol=atol)


def test_patches_bit_exact(golden):
    g = golden("patches")
    for n in (196, 36):
        out = O.extract_patches(T(g["imgs"]), T(g[f"theta{n}"]), n)
        assert out.shape == g[f"mosaic{n}"].shape
        assert torch.equal(out, T(g[f"mosaic{n}"])), n


def test_landmark_post_bit_exact(golden):
    g = golden("landmark_post")
    # global view: noise only; RNG call order of ViT_face.py:1361 on the CPU default generator
    torch.manual_seed(int(g["seed_global"]))
    noise = torch.randn(3, 196, 2) * 5
    th = O.landmark_post(T(g["raw_global"]), noise)
    assert torch.equal(th, T(g["theta_global"]))
    # local view: noise, then randint(0,196,(b,36,1))  (ViT_face.py:1361,1366)
    torch.manual_seed(int(g["seed_local"]))
    noise = torch.randn(3, 196, 2) * 5
    idx = torch.randint(0, 196, (3, 36, 1))
    th = O.landmark_post(T(g["raw_local"]), noise, idx)
    assert torch.equal(th, T(g["theta_local"]))
    th = O.landmark_post(T(g["raw_local"]))
    assert torch.equal(th, T(g["theta_plain"]))
    # joint min-max: every sample has an exact 0 and an exact 111
    assert (th.reshape(3, -1).min(1)[0] == 0).all() and (th.reshape(3, -1).max(1)[0] == 111).all()
    mos = O.extract_patches(T(g["x_aug"]).float(), T(g["theta_local"]), 36)
    assert torch.equal(mos, T(g["mosaic_local_from_half"]))


def test_tokens_and_patch_embed(golden):
    g, p = golden("patch_embed"), golden("patches")
    tok = O.tokens_from_mosaic(T(p["mosaic196"]))
    assert tok.shape == (2, 196, 192)
    assert float(tok.double().sum()) == float(g["tokens_checksum"])
    assert torch.equal(tok[0, :4], T(g["tokens_row0"]))
    y = O.patch_embed(tok, T(g["weight"]), T(g["bias"]))
    assert torch.equal(y, T(g["embedded"]))
    # layout identity tok[b,k,(i*8+j)*3+c] == mosaic[b,c,r*8+i,q*8+j]
    m = T(p["mosaic196"])
    for (b, k, i, j, c) in [(0, 0, 0, 0, 0), (1, 37, 3, 5, 2), (0, 195, 7, 7, 1)]:
        r, q = divmod(k, 14)
        assert tok[b, k, (i * 8 + j) * 3 + c] == m[b, c, r * 8 + i, q * 8 + j]


def test_dino_loss_grad_center(golden):
    for name in ("dino", "dino_bf16"):
        g = golden(name)
        s, t = T(g["student"]), T(g["teacher"])
        if name == "dino_bf16":
            s, t = s.view(torch.bfloat16).float(), t.view(torch.bfloat16).float()
        ncrops, temp = int(g["ncrops"]), float(g["temp"])
        loss, grad = O.dino_loss_and_grad(s, t, T(g["center0"]), ncrops, temp)
        assert torch.equal(loss, T(g["loss"])), name
        assert torch.equal(grad, T(g["grad_student"])), name
        c1 = O.dino_center_update(T(g["center0"]), t)
        assert torch.equal(c1, T(g["center1"])), name
    sch = O.teacher_temp_schedule(0.04, 0.07, 30, 41)
    assert np.array_equal(sch, golden("dino")["temp_schedule"])


def test_ema_bit_exact(golden):
    g = golden("ema")
    n = sum(1 for k in g if k.startswith("q"))
    q = [T(g[f"q{i}"]) for i in range(n)]
    k = [T(g[f"k0_{i}"]).clone() for i in range(n)]
    sched = O.cosine_scheduler(0.996, 1, 41, 100)
    assert np.array_equal(sched[:64], g["sched_head"])
    m = sched[int(g["it"])]
    assert m == float(g["m"])
    O.ema_update_(k, q, m)
    for i in range(n):
        assert torch.equal(k[i], T(g[f"k1_{i}"])), i


def test_cosface_hard_soft(golden):
    g = golden("cosface")
    x, w, lab = T(g["x"]), T(g["weight"]), T(g["label"])
    loss, logits, gx, gw = O.head_loss_and_grads(x, w, lab, "cosface")
    assert torch.equal(logits, T(g["logits_hard"]))
    assert torch.equal(loss, T(g["loss_hard"]))
    # GEMM reduction order depends on the host thread count -> last-ulp tolerance for grads
    close(gx, T(g["grad_x_hard"])); close(gw, T(g["grad_w_hard"]))
    lam = float(g["lam"])
    assert (g["soft_target_nnz"] <= 2).all()
    loss, logits, gx, gw = O.head_loss_and_grads(x, w, lab, "cosface", label_b=T(g["label_b"]), lam=lam)
    assert torch.equal(logits, T(g["logits_soft"]))
    assert torch.equal(loss, T(g["loss_soft"]))
    close(gx, T(g["grad_x_soft"])); close(gw, T(g["grad_w_soft"]))
    # closed form used by the fused kernels: s*(cos - m*target)   
    cos = torch.nn.functional.normalize(x) @ torch.nn.functional.normalize(w).t()
    tgt = O.mixup_target(lab, w.shape[0], lam)
    assert (64.0 * (cos - 0.4 * tgt) - logits).abs().max() < 2e-5


def test_shard_mapping(golden):
    for row in golden("shards")["table"]:
        C, R, sizes = int(row[0]), int(row[1]), [int(v) for v in row[2:] if v > 0]
        b = [hi - lo for lo, hi in O.shard_bounds(C, R) if hi > lo]
        assert b == sizes, (C, R)
        lab = np.array([0, 1, C // 2, C - 1])
        sh, loc = O.label_to_shard(lab, C, R)
        bounds = O.shard_bounds(C, R)
        for l, s_, o in zip(lab, sh, loc):
            assert bounds[s_][0] + o == l and bounds[s_][0] <= l < bounds[s_][1]


def test_arcface_unpinned_sanity():
    torch.manual_seed(0)
    x, w = torch.randn(4, 32), torch.randn(50, 32)
    lab = torch.tensor([0, 7, 49, 3])
    z = O.arcface_logits(x, w, lab)
    cos = torch.nn.functional.normalize(x) @ torch.nn.functional.normalize(w).t()
    off = torch.ones_like(cos, dtype=torch.bool)
    off[torch.arange(4), lab] = False
    assert torch.allclose(z[off], 64 * cos[off])
    th = torch.acos(cos[~off].clamp(-1, 1))
    assert torch.allclose(z[~off], 64 * torch.cos(th + 0.5), atol=1e-4)


def test_dino_head_tail_and_loss_golden(golden):
    """(f1) F.normalize + weight-normed last_layer (vision_transformer.py:296-300) feeding DINOLoss: the oracle
    against the reference's own student/teacher DINOHead + DINOLoss run (tests/golden/make_golden_dino_head.py)."""
    g = golden("dino_head")
    xs, xt, vs, gs, vt, gt = (T(g[k]) for k in ("xs", "xt", "vs", "gs", "vt", "gt"))
    close(O.dino_head_logits(xs, vs, gs)[0], T(g["student_logits_row0"]), rtol=1e-5, atol=1e-6)
    close(O.dino_head_logits(xt, vt, gt)[0], T(g["teacher_logits_row0"]), rtol=1e-5, atol=1e-6)
    loss, dx, dv, dg, c1 = O.dino_head_loss_and_grads(xs, xt, vs, gs, vt, gt, T(g["center0"]), int(g["ncrops"]),
                                                      float(g["temp"]))
    close(loss, T(g["loss"]), rtol=1e-6, atol=0)
    for a, b in ((dx, T(g["grad_xs"])), (dv, T(g["grad_vs"])), (dg.reshape(-1, 1), T(g["grad_gs"]))):
        close(a, b, rtol=1e-4, atol=1e-5 * float(b.abs().max()))      # fp32 summation order of autograd
    close(c1, T(g["center1"]), rtol=1e-6, atol=1e-7)
