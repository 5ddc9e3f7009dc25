"""CosFace / ArcFace margin heads with the reference's module surface on tcgen05 kernels.

    CosFace(in_features, out_features, device_id, s=64.0, m=0.4)     face_pre_pro/ViT_face.py:26-96
      .forward(input, label) -> [B, C] fp32 logits   (label [B] int or [B, C] float soft targets)
      .forward_loss(input, label, label_b=None, lam=1.0) -> scalar CE loss, logits never materialised
      parameter `weight` [C, D]  (registered as `loss` in the backbone -> checkpoint key loss.weight)
    ArcFace(...)  same surface; the reference names it (ViT_face.py:416-417) but never defines it,
      so its formula follows the ArcFace paper (parity unpinned, see oracle/lafs_oracle.py).

Class sharding: `shard=(rank, world)` keeps only rows [lo, hi) of torch.chunk(weight, world)
(ViT_face.py:56) on this rank; forward_loss then exchanges only per-row (max, sum-exp,
target-logit) records over the process group.
"""
import math

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401  (kept for parity of the module namespace)
from torch.nn import Parameter

from . import _lib

KIND_COSFACE, KIND_ARCFACE = 0, 1


def shard_bounds(num_classes, world):
    """[lo, hi) per rank, identical to torch.chunk(weight, world, dim=0): chunk = ceil(C/R)."""
    step = -(-num_classes // world)
    return [(min(r * step, num_classes), min((r + 1) * step, num_classes)) for r in range(world)]


def label_to_shard(label, num_classes, world):
    """(owning rank, local row) of each label -- integer, bit-exact with torch.chunk's layout."""
    step = -(-num_classes // world)
    return torch.div(label, step, rounding_mode="floor"), label % step


def two_hot_from_dense(target):
    """Recovers (label_a, label_b, lam) from a dense mixup target [B, C] (util/mixup_my.py:18-24:
    lam at label[i], 1-lam at label[B-1-i], at most two non-zeros per row, ONE lam for the whole batch).
    One device->host read (the reference builds the dense target on the host in the first place, ViT_face.py:64-71)."""
    vals, idx = target.topk(2, dim=1)
    la, lb = idx[:, 0], idx[:, 1]
    lb = torch.where(vals[:, 1] == 0, la, lb)
    # rows whose two classes coincide (or hard rows) have val0 == 1; lam is a batch scalar otherwise
    mixed = vals[:, 1] != 0
    big = torch.where(mixed, vals[:, 0], torch.full_like(vals[:, 0], -1.0))
    small = torch.where(mixed, vals[:, 0], torch.full_like(vals[:, 0], 2.0))
    too_many, any_mixed, lam_hi, lam_lo = torch.stack([
        ((target != 0).sum(1) > 2).any().to(vals.dtype), mixed.any().to(vals.dtype), big.max(), small.min()]).tolist()
    if too_many:
        raise ValueError("dense soft targets with more than two non-zeros per row are not supported "
                         "by the fused head (the reference's Mixup produces at most two)")
    if any_mixed and lam_hi - lam_lo > 1e-6:
        raise ValueError("dense soft targets with a different mixing weight per row (Mixup modes 'elem' / 'pair') are "
                         f"not supported by the fused head: weights span [{lam_lo}, {lam_hi}]; pass label_a, label_b, lam")
    lam = float(lam_hi) if any_mixed else 1.0
    return la, lb, lam


def _prep(x, want_inv=False):
    x = x.detach().contiguous()
    R, D = x.shape
    out = torch.empty(R, D, dtype=torch.bfloat16, device=x.device)
    inv = torch.empty(R, dtype=torch.float32, device=x.device) if want_inv else None
    _lib.call("lafs_normalize_rows", x.data_ptr(), _lib.dtype_code(x), R, D, out.data_ptr(), _lib.ptr(inv),
              _lib.stream())
    return out, inv


def _round8(n):
    """leading dimension of the bf16 logit gradient: a multiple of 16 elements keeps every row 32-byte
    aligned (256-bit stores in the gradient epilogue; the GEMM operands need a multiple of 8)."""
    return (n + 15) // 16 * 16


def _dw_diag_enabled(D):
    """dW with the F.normalize Jacobian's rank-one term on the tensor core (lafs_head_grad_logits_t +
    lafs_head_bwd_weight_t): no second pass over dW.  Measured on B200 (round 2): 347.5 -> 316.4 us per step at
    B=512, C=93431 and 1067 -> 999 us at B=1024, C=205990.  LAFS_DW_DIAG=0 selects GEMM + normalize_bwd pass."""
    import os
    return os.environ.get("LAFS_DW_DIAG", "1") not in ("", "0") and D % 64 == 0


def _head_backward(G, ldg, e_hat, w_hat, inv_e, inv_w, B, C, D, sharded, xchg=None, tpart=None, rows=None):
    """dE [B,D], dW [C,D] (fp32) from the bf16 logit gradient G [B, ldg] via two tcgen05 GEMMs.
    rows=(lo, n): batch-sharded caller -- only rows [lo, lo+n) of dE are wanted on this rank: the partial dE_hat of
    the class shards is reduce-scattered (NCCL) instead of all-reduced, and the Jacobian runs on those rows only."""
    dev = e_hat.device
    nbytes = _lib.lib().lafs_head_bwd_workspace_bytes(B, C, D)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    de_hat = xchg.de_in if (sharded and xchg is not None) else torch.empty(B, D, dtype=torch.float32, device=dev)
    _lib.call("lafs_head_bwd_embed", G.data_ptr(), ldg, w_hat.data_ptr(), B, C, D, de_hat.data_ptr(),
              ws.data_ptr(), nbytes, _lib.stream())
    if sharded and xchg is not None:
        de_hat = xchg.allreduce_de()                 # sum over the class shards through peer memory
        if rows is not None:
            de_hat = de_hat[rows[0]:rows[0] + rows[1]]
    elif sharded and rows is not None:
        mine = torch.empty(rows[1], D, dtype=torch.float32, device=dev)
        dist.reduce_scatter_tensor(mine, de_hat)     # sum over the class shards, each rank keeps its batch slice
        de_hat = mine
    elif sharded:
        dist.all_reduce(de_hat)                      # sum over the class shards
    e_rows, inv_rows, nb = e_hat, inv_e, B          # the rows whose dE this rank returns (dW below needs ALL rows of e_hat)
    if rows is not None:
        e_rows, inv_rows, nb = e_hat[rows[0]:rows[0] + rows[1]], inv_e[rows[0]:rows[0] + rows[1]], rows[1]
    de = torch.empty_like(de_hat)
    _lib.call("lafs_normalize_bwd", de_hat.data_ptr(), e_rows.data_ptr(), inv_rows.data_ptr(), nb, D, de.data_ptr(),
              _lib.stream())
    dw = torch.empty(C, D, dtype=torch.float32, device=dev)
    if tpart is not None:
        _lib.call("lafs_head_bwd_weight_t", G.data_ptr(), ldg, e_hat.data_ptr(), w_hat.data_ptr(), inv_w.data_ptr(),
                  tpart.data_ptr(), tpart.shape[0], tpart.shape[1], B, C, D, dw.data_ptr(), _lib.stream())
    else:
        _lib.call("lafs_head_bwd_weight", G.data_ptr(), ldg, e_hat.data_ptr(), w_hat.data_ptr(), inv_w.data_ptr(),
                  B, C, D, dw.data_ptr(), _lib.stream())
    return de, dw


class _HeadLossFn(torch.autograd.Function):
    """Fused margin head + softmax cross-entropy.  Nothing of size [B, C] is saved for backward:
    the bf16 logit gradient is recomputed on the tensor cores from the saved row statistics."""

    @staticmethod
    def forward(ctx, input, weight, head, la, lb, lam):
        ctx.rows = None
        if getattr(head, "batch_sharded", False) and head.shard is not None and head.shard[1] > 1:
            # every rank holds B/R samples; the class shards need all of them -> all_gather of the
            # embeddings and labels (the reference's only precedent is the replicated batch, ViT_face.py:55-64)
            world, rank = head.shard[1], head.shard[0]
            b_loc = input.shape[0]
            full = torch.empty(world * b_loc, input.shape[1], dtype=input.dtype, device=input.device)
            dist.all_gather_into_tensor(full, input.detach().contiguous())
            la_f = torch.empty(world * b_loc, dtype=la.dtype, device=la.device)
            dist.all_gather_into_tensor(la_f, la)
            if lb is not None:
                lb_f = torch.empty(world * b_loc, dtype=lb.dtype, device=lb.device)
                dist.all_gather_into_tensor(lb_f, lb)
                lb = lb_f
            input, la = full, la_f
            ctx.rows = (rank * b_loc, b_loc)
        B, D = input.shape
        C = weight.shape[0]
        e_hat, inv_e = _prep(input, want_inv=True)
        w_hat, inv_w = _prep(weight, want_inv=True)
        dev = input.device
        stats = torch.empty(B, 4, dtype=torch.float32, device=dev)
        nbytes = _lib.lib().lafs_head_workspace_bytes(B, C, D)
        if nbytes == 0:
            raise ValueError(f"unsupported head shape B={B} C={C} D={D}")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.call("lafs_head_fwd", e_hat.data_ptr(), w_hat.data_ptr(), la.data_ptr(), _lib.ptr(lb), lam,
                  B, C, D, head.class_lo, float(head.s), float(head.m), head.kind, stats.data_ptr(),
                  ws.data_ptr(), nbytes, _lib.stream())
        sharded = head.shard is not None and head.shard[1] > 1
        xchg = head._exchange(B, D, dev) if sharded else None
        if sharded and xchg is not None:
            stats = xchg.merge_stats(stats)          # put / flag / merge through peer memory, one kernel
        elif sharded:
            parts = torch.empty(head.shard[1], B, 4, dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(parts, stats)
            stats = torch.empty_like(stats)
            _lib.call("lafs_head_merge", parts.data_ptr(), head.shard[1], B, stats.data_ptr(), _lib.stream())
        loss = torch.empty((), dtype=torch.float32, device=dev)
        lse2 = torch.empty(B, dtype=torch.float32, device=dev)
        _lib.call("lafs_head_loss", stats.data_ptr(), la.data_ptr(), _lib.ptr(lb), lam, B, lse2.data_ptr(),
                  loss.data_ptr(), _lib.stream())
        ctx.save_for_backward(e_hat, w_hat, inv_e, inv_w, la, lb if lb is not None else la, lse2)
        ctx.cfg = (head.class_lo, float(head.s), float(head.m), head.kind, lam, lb is not None, sharded,
                   input.dtype)
        ctx.xchg = xchg
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        e_hat, w_hat, inv_e, inv_w, la, lb, lse2 = ctx.saved_tensors
        class_lo, s, m, kind, lam, has_b, sharded, in_dtype = ctx.cfg
        B, D = e_hat.shape
        C = w_hat.shape[0]
        ldg = _round8(C)
        G = torch.empty(B, ldg, dtype=torch.bfloat16, device=e_hat.device)
        g = grad_loss.detach().float().contiguous()
        tpart = None
        if _dw_diag_enabled(D):
            ldt = (C + 31) // 32 * 32
            tpart = torch.empty(4 * ((B + 127) // 128), ldt, dtype=torch.float32, device=e_hat.device)
            _lib.call("lafs_head_grad_logits_t", e_hat.data_ptr(), w_hat.data_ptr(), la.data_ptr(),
                      lb.data_ptr() if has_b else None, lam, B, C, D, class_lo, s, m, kind, lse2.data_ptr(),
                      g.data_ptr(), s / B, G.data_ptr(), ldg, tpart.data_ptr(), ldt, _lib.stream())
        else:
            _lib.call("lafs_head_grad_logits", e_hat.data_ptr(), w_hat.data_ptr(), la.data_ptr(),
                      lb.data_ptr() if has_b else None, lam, B, C, D, class_lo, s, m, kind, lse2.data_ptr(),
                      g.data_ptr(), s / B, G.data_ptr(), ldg, _lib.stream())
        de, dw = _head_backward(G, ldg, e_hat, w_hat, inv_e, inv_w, B, C, D, sharded, ctx.xchg, tpart, rows=ctx.rows)
        return de.to(in_dtype), dw, None, None, None, None


class _HeadLogitsFn(torch.autograd.Function):
    """CosFace.forward parity path: full fp32 logits out, gradients through the same GEMMs."""

    @staticmethod
    def forward(ctx, input, weight, head, la, lb, lam):
        B, D = input.shape
        C = weight.shape[0]
        e_hat, inv_e = _prep(input, want_inv=True)
        w_hat, inv_w = _prep(weight, want_inv=True)
        out = torch.empty(B, C, dtype=torch.float32, device=input.device)
        _lib.call("lafs_head_logits", e_hat.data_ptr(), w_hat.data_ptr(), la.data_ptr(), _lib.ptr(lb), lam,
                  B, C, D, head.class_lo, float(head.s), float(head.m), head.kind, out.data_ptr(), C, _lib.stream())
        ctx.save_for_backward(e_hat, w_hat, inv_e, inv_w, la)
        ctx.cfg = (float(head.s), float(head.m), head.kind, head.class_lo, input.dtype)
        return out

    @staticmethod
    def backward(ctx, grad_logits):
        e_hat, w_hat, inv_e, inv_w, la = ctx.saved_tensors
        s, m, kind, class_lo, in_dtype = ctx.cfg
        B, D = e_hat.shape
        C = w_hat.shape[0]
        ldg = _round8(C)
        G = torch.zeros(B, ldg, dtype=torch.bfloat16, device=e_hat.device)
        gz = grad_logits.detach().float() * s                      # d z / d cos = s (margin is a shift)
        if kind == KIND_ARCFACE:                                   # target column: s * d phi / d cos (no host sync)
            loc = la - class_lo
            own = (loc >= 0) & (loc < C)
            cols = loc.clamp(0, C - 1)
            rows = torch.arange(B, device=e_hat.device)
            cos_t = (e_hat.float() * w_hat[cols].float()).sum(1)
            th = math.cos(math.pi - m)
            sine = torch.sqrt((1 - cos_t * cos_t).clamp(1e-12, 1))
            dphi = torch.where(own & (cos_t > th), math.cos(m) + cos_t * math.sin(m) / sine, torch.ones_like(cos_t))
            gz[rows, cols] = gz[rows, cols] * dphi
        G[:, :C] = gz.to(torch.bfloat16)
        de, dw = _head_backward(G, ldg, e_hat, w_hat, inv_e, inv_w, B, C, D, False)
        return de.to(in_dtype), dw, None, None, None, None


class _MarginHead(nn.Module):
    kind = KIND_COSFACE

    def __init__(self, in_features, out_features, device_id, s=64.0, m=0.4, shard=None, batch_sharded=False):
        """shard=(rank, world): keep the torch.chunk class slice of this rank (ViT_face.py:56).
        batch_sharded (with shard): forward_loss takes this rank's B/world samples, all-gathers embeddings and
        labels, and its backward reduce-scatters dE back to the owners; otherwise every rank is handed
        the global batch and dE is all-reduced.  The loss is the mean over the GLOBAL batch in both forms."""
        super().__init__()
        self.batch_sharded = bool(batch_sharded)
        self.in_features = in_features
        self.out_features = out_features
        self.device_id = device_id
        self.s = s
        self.m = m
        self.shard = shard
        lo, hi = (0, out_features) if shard is None else shard_bounds(out_features, shard[1])[shard[0]]
        self.class_lo, self.class_hi = lo, hi
        full = torch.empty(out_features, in_features)
        nn.init.xavier_uniform_(full)                     # same init as the reference (ViT_face.py:46-47)
        self.weight = Parameter(full[lo:hi].clone())

    # ---- class-sharded exchange ------------------------------------------------------------------
    def enable_peer_exchange(self, group=None, enabled=True):
        """Run the two exchange steps of forward_loss (row statistics, dE sum) as peer-memory kernels over
        NVLink (csrc/exchange.cu) instead of NCCL all_gather / all_reduce.  All ranks of `group` (one node)
        must call this; buffers are created lazily per (B, D)."""
        self._peer = bool(enabled)
        self._peer_group = group
        self._xchg = {}
        return self

    def _exchange(self, B, D, dev):
        if not getattr(self, "_peer", False):
            return None
        key = (B, D)
        if key not in self._xchg:
            from .peer_exchange import PeerExchange
            self._xchg[key] = PeerExchange(self._peer_group, B, D, dev)
        return self._xchg[key]

    # ---- checkpoints of a class-sharded head ------------------------------------------------------
    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        """A full [C, D] `weight` (a reference / unsharded checkpoint, key `loss.weight`) loads into a sharded
        head by taking this rank's torch.chunk slice [class_lo, class_hi) (ViT_face.py:56)."""
        key = prefix + "weight"
        w = state_dict.get(key)
        if w is not None and w.dim() == 2 and w.shape[0] == self.out_features and self.weight.shape[0] != self.out_features:
            state_dict = dict(state_dict)
            state_dict[key] = w[self.class_lo:self.class_hi]
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    @torch.no_grad()
    def full_weight(self, group=None):
        """The full [C, D] weight on every rank of a sharded head (all-gather of the slices, in torch.chunk
        order) -- what `state_dict()` of the reference's unsharded CosFace holds; use it to save checkpoints."""
        if self.shard is None or self.shard[1] == 1:
            return self.weight.detach().clone()
        world = self.shard[1]
        step = -(-self.out_features // world)
        pad = torch.zeros(step, self.in_features, dtype=self.weight.dtype, device=self.weight.device)
        pad[: self.weight.shape[0]] = self.weight
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        return torch.cat(parts)[: self.out_features].contiguous()

    # ---- helpers -----------------------------------------------------------------------------
    # The reference's one_hot.scatter_ (ViT_face.py:66-68) raises on a label outside [0, out_features); a label the
    # kernels never meet only drops its target term (loss = lse).  Checking costs a device->host read per call, so it
    # is opt-in: head.check_labels = True.
    check_labels = False

    def _labels(self, label, label_b, lam):
        if label.dim() > 1:                                # dense [B, C] soft targets (reference API)
            if self.kind == KIND_ARCFACE:
                raise ValueError("ArcFace takes hard labels")
            la, lb, lam = two_hot_from_dense(label)
            return la.to(torch.int64).contiguous(), lb.to(torch.int64).contiguous(), lam
        la = label.to(torch.int64).contiguous()
        lb = None if label_b is None else label_b.to(torch.int64).contiguous()
        if self.check_labels:
            both = la if lb is None else torch.cat([la, lb])
            lo, hi = torch.stack([both.min(), both.max()]).tolist()
            if lo < 0 or hi >= self.out_features:
                raise IndexError(f"label out of range: [{lo}, {hi}] outside [0, {self.out_features})")
        return la, lb, float(lam)

    def _operands(self, input):
        _lib.require_cuda(input, self.weight)
        if input.dim() != 2 or input.shape[1] != self.in_features:
            raise ValueError(f"input must be [B, {self.in_features}], got {tuple(input.shape)}")
        e_hat, _ = _prep(input)
        w_hat, _ = _prep(self.weight)
        return e_hat, w_hat

    # ---- reference surface ---------------------------------------------------------------------
    def forward(self, input, label):
        """Full logits s*(cos - m*target) for the local classes (all classes when unsharded)."""
        la, lb, lam = self._labels(label, None, 1.0)
        _lib.require_cuda(input, self.weight)
        if input.dim() != 2 or input.shape[1] != self.in_features:
            raise ValueError(f"input must be [B, {self.in_features}], got {tuple(input.shape)}")
        return _HeadLogitsFn.apply(input, self.weight, self, la, lb, lam)

    def forward_stats(self, input, label, label_b=None, lam=1.0):
        """Per-row (max2, sum-exp, z_a, z_b) over this rank's classes; [B,4] fp32."""
        la, lb, lam = self._labels(label, label_b, lam)
        e_hat, w_hat = self._operands(input)
        B, C = input.shape[0], self.weight.shape[0]
        stats = torch.empty(B, 4, dtype=torch.float32, device=input.device)
        nbytes = _lib.lib().lafs_head_workspace_bytes(B, C, self.in_features)
        if nbytes == 0:
            raise ValueError(f"unsupported head shape B={B} C={C} D={self.in_features}")
        ws = torch.empty(nbytes, dtype=torch.uint8, device=input.device)
        _lib.call("lafs_head_fwd", e_hat.data_ptr(), w_hat.data_ptr(), la.data_ptr(), _lib.ptr(lb), lam,
                  B, C, self.in_features, self.class_lo, float(self.s), float(self.m), self.kind,
                  stats.data_ptr(), ws.data_ptr(), nbytes, _lib.stream())
        return stats, (la, lb, lam, e_hat, w_hat)

    def forward_loss(self, input, label, label_b=None, lam=1.0):
        """Fused head + cross-entropy, differentiable w.r.t. input and weight: the mean (over the
        batch every rank sees) of CrossEntropyLoss / SoftTargetCrossEntropy applied to
        CosFace.forward's logits (train_largescale.py:815-820), without materialising them."""
        la, lb, lam = self._labels(label, label_b, lam)
        _lib.require_cuda(input, self.weight)
        if input.dim() != 2 or input.shape[1] != self.in_features:
            raise ValueError(f"input must be [B, {self.in_features}], got {tuple(input.shape)}")
        return _HeadLossFn.apply(input, self.weight, self, la, lb, lam)

    @torch.no_grad()
    def forward_loss_stats(self, input, label, label_b=None, lam=1.0):
        """(loss, row_lse2) without autograd (evaluation / tests)."""
        stats, (la, lb, lam, _, _) = self.forward_stats(input, label, label_b, lam)
        B = input.shape[0]
        if self.shard is not None and self.shard[1] > 1:
            parts = torch.empty(self.shard[1], B, 4, dtype=torch.float32, device=input.device)
            dist.all_gather_into_tensor(parts, stats)
            merged = torch.empty_like(stats)
            _lib.call("lafs_head_merge", parts.data_ptr(), self.shard[1], B, merged.data_ptr(), _lib.stream())
            stats = merged
        loss = torch.empty((), dtype=torch.float32, device=input.device)
        lse2 = torch.empty(B, dtype=torch.float32, device=input.device)
        _lib.call("lafs_head_loss", stats.data_ptr(), la.data_ptr(), _lib.ptr(lb), lam, B, lse2.data_ptr(),
                  loss.data_ptr(), _lib.stream())
        return loss, lse2

    def __repr__(self):
        return self.__class__.__name__ + '(' \
            + 'in_features = ' + str(self.in_features) \
            + ', out_features = ' + str(self.out_features) \
            + ', s = ' + str(self.s) \
            + ', m = ' + str(self.m) + ')'


class CosFace(_MarginHead):
    kind = KIND_COSFACE

    def __init__(self, in_features, out_features, device_id, s=64.0, m=0.4, shard=None, batch_sharded=False):
        super().__init__(in_features, out_features, device_id, s, m, shard, batch_sharded)


class ArcFace(_MarginHead):
    kind = KIND_ARCFACE

    def __init__(self, in_features, out_features, device_id, s=64.0, m=0.5, shard=None, batch_sharded=False):
        super().__init__(in_features, out_features, device_id, s, m, shard, batch_sharded)
