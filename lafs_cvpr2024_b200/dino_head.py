"""DINOHead with the reference's module surface (vision_transformer.py:265-301) and the fused
last_layer -> DINOLoss path: the [(ncrops+2)B, out_dim] logits never exist in HBM.

    DINOHead(in_dim, out_dim, use_bn=False, norm_last_layer=True, nlayers=3, hidden_dim=2048,
             bottleneck_dim=256, fused_loss=False)
      .forward(x) -> logits [rows, out_dim]                       (reference behaviour, default)
      .forward(x) -> DeferredLogits(bottleneck features, head)    (fused_loss=True)
      state-dict keys mlp.*, last_layer.weight_g, last_layer.weight_v   (as the reference's checkpoints)

    DINOLoss.forward(student_output, teacher_output, epoch) accepts DeferredLogits for both arguments and then
    runs `fused_dino_loss` below: loss, centre update and the gradients w.r.t. the student's bottleneck features
    and last_layer.weight_v / weight_g, identical (to bf16-operand rounding) to
        dino_loss(student_head.last_layer(F.normalize(xs)), teacher_head.last_layer(F.normalize(xt)), epoch)
    so lafs_train.py:581-583 runs unchanged:  teacher(images[:2]) / student(images) return what their head returns
    (utils.py:635) and only dino_loss looks at it.

How (csrc/dino_head.cu): every contraction is one of the margin head's tcgen05 GEMMs --
    teacher  statistics  lafs_head_fwd          on [x_hat_t | 1 1 1 | 0] x [w_t | -c_hi -c_mid -c_lo | 0]   (centre inside the GEMM)
             Q (bf16)    lafs_head_grad_logits  (labels -1: plain softmax probabilities), U = Q . W_s  lafs_head_bwd_embed
    student  statistics  lafs_head_fwd,   loss = mean[ lse_s - <U, x_hat_s>/ts ]  (sum_k q_k = 1)
    backward P_s (bf16)  lafs_head_grad_logits, O = P_s . W_s, dW = (P_s ; Q)^T (cnt x_hat_s ; -X~)  lafs_gemm_tn
-- with streaming kernels for operand preparation (F.normalize, weight_norm, centre split, column sums), the loss,
the F.normalize / weight_norm Jacobians.  No CPU fallback.
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib

_const_cache = {}


def _minus_one_labels(n, dev):
    """int64 [-1]*n: `no target class` for the head kernels (their label test never matches)."""
    key = ("m1", dev, n)
    t = _const_cache.get(key)
    if t is None:
        t = _const_cache[key] = torch.full((n,), -1, dtype=torch.int64, device=dev)
    return t


def _one(dev):
    key = ("one", dev)
    t = _const_cache.get(key)
    if t is None:
        t = _const_cache[key] = torch.ones((), dtype=torch.float32, device=dev)
    return t


def _overlap_enabled():
    """LAFS_DH_OVERLAP=0: everything on the calling stream (measurement / debugging)."""
    import os
    return os.environ.get("LAFS_DH_OVERLAP", "1") not in ("", "0")


def _side_stream(dev):
    key = ("side", dev)
    t = _const_cache.get(key)
    if t is None:
        t = _const_cache[key] = torch.cuda.Stream(device=dev)
    return t


def _ld_probs(K):
    """row pitch of the bf16 probability matrix: a multiple of 16 elements (32-byte rows: 256-bit stores)."""
    return (K + 15) // 16 * 16


def _row_lse2(x_hat, w, rows, K, Dk, inv_temp):
    """log2-domain log-sum-exp of every row of (x_hat . w^T) * inv_temp -- one statistics GEMM, no logits."""
    dev = x_hat.device
    nbytes = _lib.lib().lafs_head_workspace_bytes(rows, K, Dk)
    if nbytes == 0:
        raise ValueError(f"unsupported fused DINO head shape rows={rows} K={K} D={Dk} (D must be a multiple of 64, <= 768)")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    stats = torch.empty(rows, 4, dtype=torch.float32, device=dev)
    _lib.call("lafs_head_fwd", x_hat.data_ptr(), w.data_ptr(), _minus_one_labels(rows, dev).data_ptr(), None, 1.0,
              rows, K, Dk, 0, float(inv_temp), 0.0, 0, stats.data_ptr(), ws.data_ptr(), nbytes, _lib.stream())
    lse2 = torch.empty(rows, dtype=torch.float32, device=dev)
    _lib.call("lafs_dh_lse2", stats.data_ptr(), rows, lse2.data_ptr(), _lib.stream())
    return lse2


def _probs(x_hat, w, rows, K, Dk, inv_temp, lse2, out, ldp):
    """out [rows, ldp] bf16 = softmax((x_hat . w^T) * inv_temp) recomputed on the tensor cores from the row lse."""
    dev = x_hat.device
    _lib.call("lafs_head_grad_logits", x_hat.data_ptr(), w.data_ptr(), _minus_one_labels(rows, dev).data_ptr(), None,
              1.0, rows, K, Dk, 0, float(inv_temp), 0.0, 0, lse2.data_ptr(), _one(dev).data_ptr(), 1.0,
              out.data_ptr(), ldp, _lib.stream())


def _probs_times(P, ldp, w, rows, K, D):
    """[rows, D] fp32 = P [rows, K] . w [K, D]  (split-K tcgen05 GEMM)."""
    dev = w.device
    nbytes = _lib.lib().lafs_head_bwd_workspace_bytes(rows, K, D)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    out = torch.empty(rows, D, dtype=torch.float32, device=dev)
    _lib.call("lafs_head_bwd_embed", P.data_ptr(), ldp, w.data_ptr(), rows, K, D, out.data_ptr(), ws.data_ptr(), nbytes,
              _lib.stream())
    return out


def dino_head_forward(xs, xt, vs, gs, vt, gt, center, ncrops, inv_ts, inv_tt, keep_for_backward=True):
    """Fused forward.  xs [ncrops*B, D] student bottleneck features (pre-normalisation), xt [2B, D] teacher's;
    vs / vt [K, D] fp32 = last_layer.weight_v, gs / gt [K] fp32 = last_layer.weight_g (None: ones);
    center [K] fp32 (the OLD centre).  Returns (loss [] fp32, colsum [K] fp32 = sum over the teacher rows of the
    teacher logits, saved) -- `saved` feeds dino_head_backward."""
    _lib.require_cuda(xs, xt, vs, vt, center, gs, gt)
    if xs.dim() != 2 or xt.dim() != 2 or xs.shape[1] != xt.shape[1]:
        raise ValueError("student / teacher features must be 2-D with the same width")
    K, D = vs.shape
    if vt.shape != (K, D) or xs.shape[1] != D or center.numel() != K:
        raise ValueError(f"shape mismatch: features {tuple(xs.shape)}, weight_v {tuple(vs.shape)} / {tuple(vt.shape)}, "
                         f"center {tuple(center.shape)}")
    if xs.shape[0] % ncrops or xt.shape[0] % 2 or xs.shape[0] // ncrops != xt.shape[0] // 2:
        raise ValueError("row counts must be ncrops*B (student) and 2*B (teacher)")
    B = xt.shape[0] // 2
    rs, rt = ncrops * B, 2 * B
    dev = xs.device
    ex = _lib.lib().lafs_dh_extra_cols()
    Dt = D + ex
    if D % 64 or Dt > 768 or B < 1 or ncrops < 2:
        raise ValueError(f"fused DINO head: bottleneck width {D} must be a multiple of 64 and <= {768 - ex}; "
                         f"B={B} >= 1, ncrops={ncrops} >= 2")
    st = _lib.stream()
    xs_c, xt_c = xs.detach().contiguous(), xt.detach().contiguous()
    vs_c, vt_c = vs.detach().float().contiguous(), vt.detach().float().contiguous()
    gs_c = None if gs is None else gs.detach().float().contiguous().view(-1)
    gt_c = None if gt is None else gt.detach().float().contiguous().view(-1)
    c = center.detach().float().contiguous().view(-1)

    # every buffer is allocated on the calling stream; the teacher's operand preparation (HBM-bound, ~60 us at the
    # reference size) then runs on a side stream under the student's preparation and statistics GEMM
    xt_hat = torch.empty(rt, Dt, dtype=torch.bfloat16, device=dev)
    xsum = torch.empty(D, dtype=torch.float32, device=dev)
    wt = torch.empty(K, Dt, dtype=torch.bfloat16, device=dev)
    colsum = torch.empty(K, dtype=torch.float32, device=dev)
    xs_hat = torch.empty(rs, D, dtype=torch.bfloat16, device=dev)
    inv_xs = torch.empty(rs, dtype=torch.float32, device=dev)
    ws = torch.empty(K, D, dtype=torch.bfloat16, device=dev)
    inv_w = torch.empty(K, dtype=torch.float32, device=dev)
    main = torch.cuda.current_stream()
    side = _side_stream(dev) if _overlap_enabled() else None

    def teacher_operands():
        # [x_hat | 1 1 1 | 0], column sums of x_hat, [w | -centre split | 0] (+ logit column sums for the centre update)
        q = _lib.stream()
        _lib.call("lafs_dh_prep_rows", xt_c.data_ptr(), _lib.dtype_code(xt_c), rt, D, Dt, xt_hat.data_ptr(), None, q)
        _lib.call("lafs_dh_xsum", xt_hat.data_ptr(), rt, D, Dt, xsum.data_ptr(), q)
        _lib.call("lafs_dh_prep_weight", vt_c.data_ptr(), _lib.ptr(gt_c), c.data_ptr(), xsum.data_ptr(), K, D, Dt,
                  wt.data_ptr(), None, colsum.data_ptr(), q)

    if side is not None:
        side.wait_stream(main)
        with torch.cuda.stream(side):
            teacher_operands()
    else:
        teacher_operands()
    # student operands, row lse of s/ts
    _lib.call("lafs_dh_prep_rows", xs_c.data_ptr(), _lib.dtype_code(xs_c), rs, D, D, xs_hat.data_ptr(), inv_xs.data_ptr(), st)
    _lib.call("lafs_dh_prep_weight", vs_c.data_ptr(), _lib.ptr(gs_c), None, None, K, D, D, ws.data_ptr(),
              inv_w.data_ptr(), None, st)
    lse2_s = _row_lse2(xs_hat, ws, rs, K, D, inv_ts)
    if side is not None:
        main.wait_stream(side)

    # teacher: row lse of (t - c)/tt, Q = softmax (bf16, the last 2B rows of the probability matrix), U = Q . W_s
    ldp = _ld_probs(K)
    P = torch.empty((rs + rt) if keep_for_backward else rt, ldp, dtype=torch.bfloat16, device=dev)
    Q = P[rs:] if keep_for_backward else P
    lse2_t = _row_lse2(xt_hat, wt, rt, K, Dt, inv_tt)
    _probs(xt_hat, wt, rt, K, Dt, inv_tt, lse2_t, Q, ldp)
    U = _probs_times(Q, ldp, ws, rt, K, D)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    sample_loss = torch.empty(B, dtype=torch.float32, device=dev)
    _lib.call("lafs_dh_loss", lse2_s.data_ptr(), U.data_ptr(), xs_hat.data_ptr(), B, ncrops, D, float(inv_ts),
              sample_loss.data_ptr(), loss.data_ptr(), st)
    saved = dict(B=B, K=K, D=D, ncrops=ncrops, inv_ts=float(inv_ts), ldp=ldp, xs_hat=xs_hat, inv_xs=inv_xs, ws=ws,
                 inv_w=inv_w, vs=vs_c, gs=gs_c, lse2_s=lse2_s, U=U, P=P if keep_for_backward else None)
    return loss, colsum, saved


def dino_head_backward(saved, grad_out, want_grad_g=False):
    """Gradients of the fused loss: (d/d xs [ncrops*B, D] fp32, d/d weight_v [K, D] fp32, d/d weight_g [K] or None).
    grad_out: device fp32 scalar (upstream gradient of the loss, e.g. the AMP loss scale)."""
    B, K, D, ncrops, inv_ts, ldp = (saved[k] for k in ("B", "K", "D", "ncrops", "inv_ts", "ldp"))
    P = saved["P"]
    if P is None:
        raise RuntimeError("dino_head_forward ran with keep_for_backward=False")
    xs_hat, ws, U = saved["xs_hat"], saved["ws"], saved["U"]
    dev = xs_hat.device
    rs, rt = ncrops * B, 2 * B
    st = _lib.stream()
    g = grad_out.detach().float().contiguous()
    _probs(xs_hat, ws, rs, K, D, inv_ts, saved["lse2_s"], P, ldp)          # P_s into the first ncrops*B rows
    O = _probs_times(P, ldp, ws, rs, K, D)
    dx = torch.empty(rs, D, dtype=torch.float32, device=dev)
    y = torch.empty(rs + rt, D, dtype=torch.bfloat16, device=dev)
    _lib.call("lafs_dh_bwd_rows", O.data_ptr(), U.data_ptr(), xs_hat.data_ptr(), saved["inv_xs"].data_ptr(), g.data_ptr(),
              B, ncrops, D, inv_ts, dx.data_ptr(), y.data_ptr(), st)
    dv = torch.empty(K, D, dtype=torch.float32, device=dev)
    _lib.call("lafs_gemm_tn", P.data_ptr(), ldp, y.data_ptr(), D, rs + rt, K, D, dv.data_ptr(), D, st)
    dg = torch.empty(K, dtype=torch.float32, device=dev) if want_grad_g else None
    coef = inv_ts / float((2 * ncrops - 2) * B)
    _lib.call("lafs_dh_wn_bwd", dv.data_ptr(), saved["vs"].data_ptr(), _lib.ptr(saved["gs"]), saved["inv_w"].data_ptr(),
              g.data_ptr(), K, D, coef, dv.data_ptr(), _lib.ptr(dg), st)
    return dx, dv, dg


class _DinoHeadLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xs, vs, gs, xt, vt, gt, center, ncrops, inv_ts, inv_tt, stash):
        need_bwd = any(ctx.needs_input_grad[:3])
        loss, colsum, saved = dino_head_forward(xs, xt, vs, gs, vt, gt, center, ncrops, inv_ts, inv_tt,
                                                keep_for_backward=need_bwd)
        ctx.saved = saved
        ctx.meta = (xs.dtype, vs.dtype, gs.dtype, tuple(gs.shape), ctx.needs_input_grad[2])
        stash["colsum"] = colsum
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        xs_dtype, vs_dtype, gs_dtype, gs_shape, want_g = ctx.meta
        dx, dv, dg = dino_head_backward(ctx.saved, grad_loss, want_grad_g=want_g)
        ctx.saved = None                     # frees the probability matrix
        if dg is not None:
            dg = dg.view(gs_shape).to(gs_dtype)
        return dx.to(xs_dtype), dv.to(vs_dtype), dg, None, None, None, None, None, None, None, None


class DeferredLogits:
    """What a DINOHead with fused_loss=True returns instead of the [rows, out_dim] logits: the bottleneck features
    (mlp output, before F.normalize) and the head that owns the prototypes.  DINOLoss consumes it; `.logits()`
    materialises the reference's tensor (stock F.normalize + weight-normed Linear) for anything else."""

    def __init__(self, features, head):
        self.features = features
        self.head = head

    @property
    def shape(self):
        return (self.features.shape[0], self.head.last_layer.weight_v.shape[0])

    def __len__(self):
        return self.features.shape[0]

    def logits(self):
        return self.head.last_layer(nn.functional.normalize(self.features, dim=-1, p=2))


def _trunc_normal_(tensor, std):
    # utils.trunc_normal_(m.weight, std=.02) (utils.py:512-550): truncation at +-2 (absolute), far outside 0.02's tail
    return nn.init.trunc_normal_(tensor, mean=0.0, std=std, a=-2.0, b=2.0)


class DINOHead(nn.Module):
    def __init__(self, in_dim, out_dim, use_bn=False, norm_last_layer=True, nlayers=3, hidden_dim=2048,
                 bottleneck_dim=256, fused_loss=False):
        super().__init__()
        nlayers = max(nlayers, 1)
        if nlayers == 1:
            self.mlp = nn.Linear(in_dim, bottleneck_dim)
        else:
            layers = [nn.Linear(in_dim, hidden_dim)]
            if use_bn:
                layers.append(nn.BatchNorm1d(hidden_dim))
            layers.append(nn.GELU())
            for _ in range(nlayers - 2):
                layers.append(nn.Linear(hidden_dim, hidden_dim))
                if use_bn:
                    layers.append(nn.BatchNorm1d(hidden_dim))
                layers.append(nn.GELU())
            layers.append(nn.Linear(hidden_dim, bottleneck_dim))
            self.mlp = nn.Sequential(*layers)
        self.apply(self._init_weights)
        # same (deprecated) parametrisation call as the reference: keeps the checkpoint keys weight_g / weight_v
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            self.last_layer = nn.utils.weight_norm(nn.Linear(bottleneck_dim, out_dim, bias=False))
        self.last_layer.weight_g.data.fill_(1)
        if norm_last_layer:
            self.last_layer.weight_g.requires_grad = False
        self.fused_loss = bool(fused_loss)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            _trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    def forward_bottleneck(self, x):
        """mlp output: what F.normalize + last_layer are applied to (vision_transformer.py:298)."""
        return self.mlp(x)

    def forward(self, x):
        x = self.mlp(x)
        if self.fused_loss:
            return DeferredLogits(x, self)
        x = nn.functional.normalize(x, dim=-1, p=2)
        return self.last_layer(x)


def fused_dino_loss(dino_loss, student_out, teacher_out, epoch):
    """DINOLoss.forward on DeferredLogits (called by DINOLoss.forward).  Updates dino_loss.center like
    DINOLoss.update_center (the loss uses the old centre, lafs_train.py:652,666)."""
    import torch.distributed as dist
    sh, th = student_out.head, teacher_out.head
    xs, xt = student_out.features, teacher_out.features.detach()
    if xs.dim() != 2 or xt.dim() != 2:
        raise ValueError("bottleneck features must be 2-D [rows, bottleneck_dim]")
    temp = float(dino_loss.teacher_temp_schedule[epoch])
    stash = {}
    loss = _DinoHeadLossFn.apply(xs, sh.last_layer.weight_v, sh.last_layer.weight_g, xt,
                                 th.last_layer.weight_v.detach(), th.last_layer.weight_g.detach(),
                                 dino_loss.center, dino_loss.ncrops, 1.0 / dino_loss.student_temp, 1.0 / temp, stash)
    with torch.no_grad():
        colsum = stash["colsum"]
        world = 1
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size()
            if world > 1:
                colsum = dino_loss._allreduce_colsum(colsum)
        K = colsum.numel()
        center = dino_loss.center.float().contiguous()
        new_center = torch.empty_like(center)
        m = float(dino_loss.center_momentum)
        _lib.call("lafs_center_ema", center.data_ptr(), colsum.data_ptr(), float(xt.shape[0] * world),
                  float(np.float32(m)), float(np.float32(1.0 - m)), K, new_center.data_ptr(), _lib.stream())
        dino_loss.center = new_center      # re-bound, like the reference (lafs_train.py:679)
    return loss
