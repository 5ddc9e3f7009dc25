"""CPU oracle for the LAFS hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU restatement (torch-CPU / numpy, fp32 unless stated) of the reference's per-step
hot path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product package
(lafs_cvpr2024_b200) never does and has no CPU fallback.

Parity status ("pinned" = checked against the reference itself, run in the build
container through oracle/ref_harness.py, and against the committed vectors under
tests/golden/ that tests/golden/make_golden.py generated from the reference):

  extract_patches / tokens_from_mosaic   pinned   face_pre_pro/ViT_face.py:1615-1656, lafs_train.py:538
  landmark_post                          pinned   face_pre_pro/ViT_face.py:1347-1378
  patch_embed                            pinned   face_pre_pro/ViT_face.py:759-761
  dino_loss / dino_center_update         pinned   lafs_train.py:643-679
  dino_head_logits / dino_head_loss      pinned   vision_transformer.py:296-300 (F.normalize + weight-normed last_layer)
                                                  feeding lafs_train.py:643-679 (golden dino_head.npz + live test)
  ema_update                             pinned   lafs_train.py:610-613
  clip_gradients_ / student_update_      pinned   utils.py:132-141, lafs_train.py:511-517,601-613 (torch.optim.AdamW
                                                  is PyTorch itself; the clip is checked against utils.clip_gradients)
  cosface_logits                         pinned   face_pre_pro/ViT_face.py:49-89
  shard_bounds / label_to_shard          pinned   face_pre_pro/ViT_face.py:56 (torch.chunk)
  cross_entropy (hard labels)            pinned   train_largescale.py:604 (torch.nn.CrossEntropyLoss)
  soft_target_cross_entropy              PARITY UNPINNED  timm.loss.SoftTargetCrossEntropy is an
                                         un-vendored, un-pinned third-party dependency
                                         (train_largescale.py:47,602); restated from its published
                                         one-line definition.
  arcface_logits                         PARITY UNPINNED  the reference names ArcFace
                                         (face_pre_pro/ViT_face.py:416-417,654-655) but never
                                         defines it; restated from the ArcFace paper /
                                         face.evoLVe metrics.py form the CosFace class was taken from.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

PATCH = 8  # ViT_face.py:606,1264  patch_shape = tensor([8, 8])


# --------------------------------------------------------------------------------------
# (1) landmark post-processing, patch extraction, token layout, patch embedding
# --------------------------------------------------------------------------------------
def landmark_post(raw, noise=None, extract_id=None, scale=111.0):
    """Joint min-max scaling of the 392 regressed numbers, optional additive noise and
    optional landmark re-sampling.  Follows face_pre_pro/ViT_face.py:1347-1378
    (twin without noise/gather: :694-705).

    raw        [B, 2n] fp32   output of output_layer
    noise      [B, n, 2] fp32 or None   (reference: torch.randn(theta.shape)*5, :1361)
    extract_id [B, k] int64 or None     (reference: torch.randint(0,c,(b,k,1)), :1366-1370)
    returns theta [B, n or k, 2] fp32
    """
    t_max = raw.max(dim=1, keepdim=True)[0]
    t_min = raw.min(dim=1, keepdim=True)[0]
    theta = (raw - t_min) / (t_max - t_min) * scale
    theta = theta.view(raw.shape[0], -1, 2)
    if noise is not None:
        theta = theta + noise
    if extract_id is not None:
        idx = extract_id.view(raw.shape[0], -1, 1).repeat(1, 1, 2)
        theta = torch.gather(theta, 1, idx)
    return theta


def extract_patches(imgs, landmarks, num_landm=None, patch=PATCH, recip_mul=False):
    """Bilinear landmark patch mosaic.  Follows face_pre_pro/ViT_face.py:1615-1656 with the
    per-landmark python loop replaced by one grid_sample over all landmarks (the
    per-element op sequence -- add, divide by img_shape*0.5, subtract 1, then
    F.grid_sample(bilinear, zeros, align_corners=False) -- is unchanged, so results are
    bit-identical to the loop; pinned in tests/test_oracle_vs_reference.py).

    imgs [B,C,H,W] fp32, landmarks [B,n,2] fp32 (x, y) in pixels.
    Output pixel (a, b) of patch k samples x = theta_x + a - 4 - 0.5, y = theta_y + b - 4 - 0.5
    (output rows step x: SURVEY Q2).  Returns the mosaic [B, C, 8r, 8r], r = sqrt(n).

    recip_mul=True restates what eager CUDA does for `tensor / python_scalar`
    (multiply by fp32(1/56)); the CPU reference divides.  Default follows the CPU reference.
    """
    B, C, H, W = imgs.shape
    n = landmarks.shape[1] if num_landm is None else num_landm
    landmarks = landmarks[:, :n]
    half = patch / 2
    ar = torch.arange(-half, half, dtype=torch.float32, device=landmarks.device)
    # sampling_grid[a][b] = (a-4, b-4)  (ViT_face.py:1637-1640)
    sg = torch.stack(torch.meshgrid(ar, ar, indexing="ij"), dim=-1)  # [8,8,2]
    pts = sg[None, None] + landmarks[:, :, None, None, :]  # [B,n,8,8,2]
    if recip_mul:
        grid = pts * torch.tensor(1.0 / (H * 0.5), dtype=torch.float32, device=pts.device) - 1
    else:
        grid = pts / (H * 0.5) - 1
    out = F.grid_sample(imgs, grid.reshape(B, n * patch, patch, 2), align_corners=False)
    out = out.reshape(B, C, n, patch, patch)
    r = int(math.isqrt(n))
    out = out.reshape(B, C, r, r, patch, patch).permute(0, 1, 2, 4, 3, 5)
    return out.reshape(B, C, r * patch, r * patch)


def tokens_from_mosaic(mosaic, patch=PATCH):
    """einops 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)'  (lafs_train.py:538, ViT_face.py:760).
    Feature index = (p1*8 + p2)*C + c  (SURVEY Q3)."""
    B, C, HH, WW = mosaic.shape
    h, w = HH // patch, WW // patch
    x = mosaic.reshape(B, C, h, patch, w, patch).permute(0, 2, 4, 3, 5, 1)
    return x.reshape(B, h * w, patch * patch * C)


def extract_tokens(imgs, landmarks, num_landm=None, recip_mul=False):
    return tokens_from_mosaic(extract_patches(imgs, landmarks, num_landm, recip_mul=recip_mul))


def patch_embed(tokens, weight, bias):
    """patch_to_embedding = nn.Linear(192, dim)  (ViT_face.py:761)."""
    return F.linear(tokens, weight, bias)


def gather_embed(imgs, landmarks, weight, bias, round_bf16=False, recip_mul=False):
    """extract -> rearrange -> patch_to_embedding.  round_bf16=True evaluates the fp32
    math on bf16-rounded tokens / weights (what a bf16 tensor-core path sees: SURVEY H4)."""
    tok = extract_tokens(imgs, landmarks, recip_mul=recip_mul)
    w = weight
    if round_bf16:
        tok = tok.bfloat16().float()
        w = weight.bfloat16().float()
    return F.linear(tok, w, bias)


# --------------------------------------------------------------------------------------
# (2) DINO loss + centre update
# --------------------------------------------------------------------------------------
def teacher_temp_schedule(warmup_teacher_temp, teacher_temp, warmup_epochs, nepochs):
    """lafs_train.py:637-641."""
    return np.concatenate((np.linspace(warmup_teacher_temp, teacher_temp, warmup_epochs),
                           np.ones(nepochs - warmup_epochs) * teacher_temp))


def dino_loss(student_output, teacher_output, center, ncrops, teacher_temp, student_temp=0.1):
    """DINOLoss.forward without the centre side effect.  Follows lafs_train.py:643-667.
    student_output [ncrops*B, K], teacher_output [2B, K], center [1,K]; fp32 math."""
    student_out = (student_output.float() / student_temp).chunk(ncrops)
    q_all = F.softmax((teacher_output.float() - center) / teacher_temp, dim=-1).chunk(2)
    total, n_terms = 0.0, 0
    for iq, q in enumerate(q_all):
        for v in range(ncrops):
            if v == iq:
                continue
            loss = torch.sum(-q * F.log_softmax(student_out[v], dim=-1), dim=-1)
            total = total + loss.mean()
            n_terms += 1
    return total / n_terms


def dino_loss_and_grad(student_output, teacher_output, center, ncrops, teacher_temp,
                       student_temp=0.1, grad_out=1.0):
    s = student_output.detach().float().clone().requires_grad_(True)
    loss = dino_loss(s, teacher_output, center, ncrops, teacher_temp, student_temp)
    (g,) = torch.autograd.grad(loss, s, torch.tensor(grad_out, dtype=loss.dtype))
    return loss.detach(), g


def dino_center_update(center, teacher_output, momentum=0.9, world_size=1, allreduced_sum=None):
    """DINOLoss.update_center, lafs_train.py:669-679.  `allreduced_sum` stands in for the
    result of dist.all_reduce over ranks (world_size>1)."""
    batch_center = torch.sum(teacher_output.float(), dim=0, keepdim=True)
    if allreduced_sum is not None:
        batch_center = allreduced_sum
    batch_center = batch_center / (len(teacher_output) * world_size)
    return center * momentum + batch_center * (1 - momentum)


# --------------------------------------------------------------------------------------
# (f1) DINOHead tail (F.normalize + weight-normed last_layer) feeding the DINO loss
# --------------------------------------------------------------------------------------
def _bf16_st(t):
    """round to bf16 with a straight-through gradient (the operand rounding of the tensor-core path)."""
    return t + (t.detach().bfloat16().float() - t.detach())


def dino_head_logits(x, weight_v, weight_g, round_bf16=False):
    """DINOHead.forward after the mlp, vision_transformer.py:298-300:
    x = F.normalize(x, dim=-1, p=2); x = last_layer(x) with last_layer = weight_norm(Linear(D, K, bias=False)),
    i.e. weight = v * (g / ||v||_row) (torch._weight_norm, dim=0).  weight_g [K,1] or [K]."""
    xh = F.normalize(x.float(), dim=-1, p=2)
    w = weight_v.float() * (weight_g.float().reshape(-1, 1) / weight_v.float().norm(dim=1, keepdim=True))
    if round_bf16:
        xh, w = _bf16_st(xh), _bf16_st(w)
    return xh @ w.t()


def dino_head_loss(xs, xt, vs, gs, vt, gt, center, ncrops, teacher_temp, student_temp=0.1, round_bf16=False):
    """loss of lafs_train.py:581-583 from the bottleneck features: student / teacher logits through their own
    last layers, then DINOLoss.forward.  Returns (loss, teacher_logits)."""
    s = dino_head_logits(xs, vs, gs, round_bf16)
    with torch.no_grad():
        t = dino_head_logits(xt, vt, gt, round_bf16)
    return dino_loss(s, t, center, ncrops, teacher_temp, student_temp), t


def dino_head_loss_and_grads(xs, xt, vs, gs, vt, gt, center, ncrops, teacher_temp, student_temp=0.1,
                             round_bf16=False, grad_out=1.0):
    """(loss, d/d xs, d/d weight_v, d/d weight_g, new centre) by autograd through the restatement."""
    xs_ = xs.detach().float().clone().requires_grad_(True)
    vs_ = vs.detach().float().clone().requires_grad_(True)
    gs_ = gs.detach().float().clone().requires_grad_(True)
    loss, t = dino_head_loss(xs_, xt, vs_, gs_, vt, gt, center, ncrops, teacher_temp, student_temp, round_bf16)
    dx, dv, dg = torch.autograd.grad(loss, (xs_, vs_, gs_), torch.tensor(grad_out, dtype=loss.dtype, device=loss.device))
    return loss.detach(), dx, dv, dg, dino_center_update(center.reshape(1, -1), t)


# --------------------------------------------------------------------------------------
# (3) teacher EMA
# --------------------------------------------------------------------------------------
def ema_update_(teacher_params, student_params, m):
    """lafs_train.py:610-613: param_k.mul_(m).add_((1 - m) * param_q).  In place."""
    with torch.no_grad():
        for q, k in zip(student_params, teacher_params):
            k.mul_(m).add_((1 - m) * q.detach())


def clip_gradients_(grads, clip):
    """utils.clip_gradients, utils.py:132-141, on a list of gradient tensors (None = no gradient): per-TENSOR
    L2 norm, `clip_coef = clip / (norm + 1e-6)`, gradient scaled in place when clip_coef < 1.  Returns the norms."""
    norms = []
    for g in grads:
        if g is not None:
            param_norm = g.norm(2)
            norms.append(param_norm.item())
            clip_coef = clip / (param_norm + 1e-6)
            if clip_coef < 1:
                g.mul_(clip_coef)
    return norms


def student_update_(params, grads, teacher, optimizer, lr, weight_decay, clip, ema_m):
    """One student update exactly as lafs_train.py:511-517,601-613 runs it: schedule values into the param groups
    (weight decay only into group 0), utils.clip_gradients, optimizer.step() (torch.optim.AdamW built over
    utils.get_params_groups' two groups), then the teacher EMA loop.  `optimizer` is a torch.optim.AdamW whose
    params are `params`; grads[i] None = cancelled / absent gradient.  In place; returns the gradient norms."""
    for i, group in enumerate(optimizer.param_groups):
        group["lr"] = lr
        if i == 0:
            group["weight_decay"] = weight_decay
    for p, g in zip(params, grads):
        p.grad = None if g is None else g.clone()
    norms = clip_gradients_([p.grad for p in params], clip) if clip else []
    optimizer.step()
    if teacher is not None:
        ema_update_(teacher, params, ema_m)
    return norms


def cosine_scheduler(base_value, final_value, epochs, niter_per_ep):
    """utils.py:187-198 without warm-up (momentum_schedule, lafs_train.py:423-424)."""
    iters = np.arange(epochs * niter_per_ep)
    return final_value + 0.5 * (base_value - final_value) * (1 + np.cos(np.pi * iters / len(iters)))


# --------------------------------------------------------------------------------------
# (4) margin heads + cross entropy + class sharding
# --------------------------------------------------------------------------------------
def _round_st(v):
    """bf16 rounding in the forward value, identity in the backward pass (straight-through)."""
    return v + (v.detach().bfloat16().float() - v.detach())


def _cosine(x, weight, pre_normalized):
    if pre_normalized == "bf16_st":
        # F.normalize, then the unit rows rounded to bf16 exactly as a bf16 tensor-core path stores them;
        # gradients still flow through the normalisation Jacobians (same operands, fp32 math)
        return F.linear(_round_st(F.normalize(x)), _round_st(F.normalize(weight)))
    if pre_normalized:   # operands already unit rows (e.g. the bf16-rounded rows a tensor core sees)
        return F.linear(x, weight)
    return F.linear(F.normalize(x), F.normalize(weight))


def cosface_logits(x, weight, label, s=64.0, m=0.4, pre_normalized=False):
    """CosFace.forward, face_pre_pro/ViT_face.py:49-89 (device_id=None branch).
    label [B] integer or [B,C] float soft targets."""
    cosine = _cosine(x, weight, pre_normalized)
    phi = cosine - m
    if label.dim() > 1:
        one_hot = label
    else:
        one_hot = torch.zeros_like(cosine)
        one_hot.scatter_(1, label.view(-1, 1).long(), 1)
    return (one_hot * phi + (1.0 - one_hot) * cosine) * s


def arcface_logits(x, weight, label, s=64.0, m=0.5, pre_normalized=False):
    """PARITY UNPINNED (see module docstring): ArcFace with the 'easy_margin=False' form."""
    cosine = _cosine(x, weight, pre_normalized)
    sine = torch.sqrt((1.0 - cosine * cosine).clamp(0, 1))
    cos_m, sin_m = math.cos(m), math.sin(m)
    th, mm = math.cos(math.pi - m), math.sin(math.pi - m) * m
    phi = cosine * cos_m - sine * sin_m
    phi = torch.where(cosine > th, phi, cosine - mm)
    one_hot = torch.zeros_like(cosine)
    one_hot.scatter_(1, label.view(-1, 1).long(), 1)
    return (one_hot * phi + (1.0 - one_hot) * cosine) * s


def cross_entropy(logits, label):
    """nn.CrossEntropyLoss(), train_largescale.py:604."""
    return F.cross_entropy(logits, label.long())


def soft_target_cross_entropy(logits, target):
    """PARITY UNPINNED: timm.loss.SoftTargetCrossEntropy (train_largescale.py:602):
    sum(-target * log_softmax(x, -1), -1).mean()."""
    return torch.sum(-target * F.log_softmax(logits, dim=-1), dim=-1).mean()


def mixup_target(label, num_classes, lam):
    """util/mixup_my.py:13-24 with smoothing=0: lam*onehot(y) + (1-lam)*onehot(y.flip(0))."""
    y1 = F.one_hot(label.long(), num_classes).float()
    y2 = F.one_hot(label.flip(0).long(), num_classes).float()
    return y1 * lam + y2 * (1.0 - lam)


def head_loss_and_grads(x, weight, label, kind="cosface", s=64.0, m=None, label_b=None, lam=1.0,
                        grad_out=1.0, pre_normalized=False):
    """logits -> CE (hard, or two-hot soft targets when label_b is given) with autograd
    gradients w.r.t. x and weight.  fp32."""
    x = x.detach().float().clone().requires_grad_(True)
    w = weight.detach().float().clone().requires_grad_(True)
    C = w.shape[0]
    if kind == "cosface":
        mm = 0.4 if m is None else m
        if label_b is None:
            logits = cosface_logits(x, w, label, s, mm, pre_normalized)
            loss = cross_entropy(logits, label)
        else:
            tgt = (F.one_hot(label.long(), C).float() * lam
                   + F.one_hot(label_b.long(), C).float() * (1.0 - lam))
            logits = cosface_logits(x, w, tgt, s, mm, pre_normalized)
            loss = soft_target_cross_entropy(logits, tgt)
    else:
        mm = 0.5 if m is None else m
        logits = arcface_logits(x, w, label, s, mm, pre_normalized)
        loss = cross_entropy(logits, label)
    gx, gw = torch.autograd.grad(loss, (x, w), torch.tensor(grad_out))
    return loss.detach(), logits.detach(), gx, gw


def shard_bounds(num_classes, world):
    """Class ranges of torch.chunk(weight, world, dim=0) (ViT_face.py:56): chunk size
    ceil(C/R), trailing chunk short (possibly fewer than `world` chunks)."""
    step = -(-num_classes // world)
    return [(min(r * step, num_classes), min((r + 1) * step, num_classes)) for r in range(world)]


def label_to_shard(label, num_classes, world):
    """shard = label // ceil(C/R); local = label % ceil(C/R)  -- bit-exact integer mapping."""
    step = -(-num_classes // world)
    label = np.asarray(label, dtype=np.int64)
    return label // step, label % step


def softmax_stats(logits):
    """Per-row (max, sum-exp) of a logit block -- the record a class shard contributes."""
    m = logits.max(dim=1)[0]
    return m, torch.exp(logits - m[:, None]).sum(1)


def merge_softmax_stats(ms, ls):
    """Merges per-shard (max, sum-exp) records into the full-row log-sum-exp: the exchange the
    class-parallel head performs instead of moving logits."""
    m = torch.stack(ms).max(0)[0]
    l = sum(li * torch.exp(mi - m) for mi, li in zip(ms, ls))
    return m + torch.log(l)
