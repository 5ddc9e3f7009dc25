"""DINOLoss with the reference's module surface (lafs_train.py:626-679) on fused sm_100a kernels.

    DINOLoss(out_dim, ncrops, warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs,
             nepochs, student_temp=0.1, center_momentum=0.9)
    .forward(student_output, teacher_output, epoch) -> scalar loss
    .update_center(teacher_output)                  (no_grad; all-reduce over the process group)
    buffer `center` [1, out_dim] fp32  (the only state-dict entry)

Forward is one streaming pass over the logits that also yields the teacher column sums, so
update_center costs one [K] all-reduce plus a [K] kernel.  As in the reference, the loss uses
the OLD centre and `self.center` is re-bound to a new tensor afterwards.
"""
import numpy as np
import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib


def _workspace(dev, nbytes):
    return torch.empty(nbytes, dtype=torch.uint8, device=dev)


class _DinoLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, student, teacher, center, ncrops, inv_ts, inv_tt, stash, momentum):
        _lib.require_cuda(student, teacher, center)
        if student.dim() != 2 or teacher.dim() != 2:
            raise ValueError("student_output / teacher_output must be 2-D [rows, out_dim]")
        K = student.shape[1]
        if student.shape[0] % ncrops or teacher.shape[0] % 2:
            raise ValueError("row counts must be ncrops*B (student) and 2*B (teacher)")
        B = student.shape[0] // ncrops
        if teacher.shape != (2 * B, K) or center.numel() != K:
            raise ValueError(f"shape mismatch: student {tuple(student.shape)}, teacher {tuple(teacher.shape)}, "
                             f"center {tuple(center.shape)}")
        s = student.detach().contiguous()
        t = teacher.detach().to(s.dtype).contiguous()
        c = center.detach().float().contiguous()
        dev = s.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        row_stats = torch.empty((ncrops + 2) * B, dtype=torch.float32, device=dev)
        colsum = torch.empty(K, dtype=torch.float32, device=dev)
        nbytes = _lib.lib().lafs_dino_workspace_bytes(B, K, ncrops)
        if nbytes == 0:
            raise ValueError(f"unsupported DINO shape B={B} K={K} ncrops={ncrops}")
        ws = _workspace(dev, nbytes)
        # one process: the centre EMA is produced by the same launch (no all-reduce in between)
        new_center = None
        if momentum is not None:
            new_center = torch.empty(1, K, dtype=torch.float32, device=dev)
        mom = float(np.float32(momentum)) if momentum is not None else 0.0
        om = float(np.float32(1.0 - momentum)) if momentum is not None else 0.0
        _lib.call("lafs_dino_fwd", s.data_ptr(), t.data_ptr(), c.data_ptr(), B, K, ncrops,
                  float(inv_ts), float(inv_tt), _lib.dtype_code(s), loss.data_ptr(),
                  row_stats.data_ptr(), colsum.data_ptr(), ws.data_ptr(), nbytes, _lib.ptr(new_center), mom, om,
                  _lib.stream())
        ctx.save_for_backward(s, t, c, row_stats)
        ctx.meta = (B, K, ncrops, float(inv_ts), float(inv_tt))
        if stash is not None:
            stash["new_center"] = new_center
            stash["colsum"] = colsum
            stash["teacher"] = (teacher.data_ptr(), teacher._version, tuple(teacher.shape))
        ctx.mark_non_differentiable(colsum)
        return loss, colsum

    @staticmethod
    def backward(ctx, grad_loss, _grad_colsum):
        s, t, c, row_stats = ctx.saved_tensors
        B, K, ncrops, inv_ts, inv_tt = ctx.meta
        g = grad_loss.detach().float().contiguous()
        grad_s = torch.empty_like(s)
        _lib.call("lafs_dino_bwd", s.data_ptr(), t.data_ptr(), c.data_ptr(), row_stats.data_ptr(),
                  g.data_ptr(), B, K, ncrops, inv_ts, inv_tt, _lib.dtype_code(s), grad_s.data_ptr(),
                  _lib.stream())
        return grad_s, None, None, None, None, None, None, None


class DINOLoss(nn.Module):
    def __init__(self, out_dim, ncrops, warmup_teacher_temp, teacher_temp,
                 warmup_teacher_temp_epochs, nepochs, student_temp=0.1,
                 center_momentum=0.9):
        super().__init__()
        self.student_temp = student_temp
        self.center_momentum = center_momentum
        self.ncrops = ncrops
        self.register_buffer("center", torch.zeros(1, out_dim))
        # same schedule object as the reference (lafs_train.py:637-641), indexed by epoch
        self.teacher_temp_schedule = np.concatenate((
            np.linspace(warmup_teacher_temp, teacher_temp, warmup_teacher_temp_epochs),
            np.ones(nepochs - warmup_teacher_temp_epochs) * teacher_temp
        ))
        self._stash = {}

    def enable_peer_exchange(self, group=None, enabled=True):
        """Sum the [K] teacher column sums over the ranks with the NVLink peer-memory all-reduce kernel
        (csrc/exchange.cu) instead of NCCL all_reduce (lafs_train.py:675).  All ranks of `group` (one node)
        must call this."""
        self._peer = bool(enabled)
        self._peer_group = group
        self._xchg = None
        return self

    def _exchange(self, K, dev):
        if not getattr(self, "_peer", False):
            return None
        if self._xchg is None:
            from .peer_exchange import PeerExchange
            self._xchg = PeerExchange(self._peer_group, 1, K, dev)
        return self._xchg

    def _allreduce_colsum(self, colsum):
        """sum of `colsum` [K] over the ranks (every rank gets identical bits)."""
        xchg = self._exchange(colsum.numel(), colsum.device)
        if xchg is None:
            dist.all_reduce(colsum)
            return colsum
        if colsum.data_ptr() != xchg.de_in.data_ptr():
            xchg.de_in.view(-1).copy_(colsum)
        return xchg.allreduce_de().view(-1)

    def forward(self, student_output, teacher_output, epoch):
        from .dino_head import DeferredLogits, fused_dino_loss
        if isinstance(student_output, DeferredLogits) or isinstance(teacher_output, DeferredLogits):
            # heads built with fused_loss=True hand over bottleneck features: last_layer GEMM + loss in one path
            if not (isinstance(student_output, DeferredLogits) and isinstance(teacher_output, DeferredLogits)):
                raise TypeError("student_output and teacher_output must both be DeferredLogits (fused_loss=True on "
                                "both heads) or both be logit tensors")
            return fused_dino_loss(self, student_output, teacher_output, epoch)
        temp = self.teacher_temp_schedule[epoch]
        single = not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)
        loss, _ = _DinoLossFn.apply(student_output, teacher_output, self.center, self.ncrops,
                                    1.0 / self.student_temp, 1.0 / float(temp), self._stash,
                                    float(self.center_momentum) if single else None)
        self.update_center(teacher_output)
        return loss

    @torch.no_grad()
    def loss_and_grad(self, student_output, teacher_output, epoch, grad_scale=None):
        """loss and d(loss)/d(student_output) in one call (training loops that own the backward of
        the head, e.g. SSLHotPath): one C call, no autograd graph.  grad_scale: optional device scalar the
        gradient is multiplied with (AMP loss scale); default 1.  Also updates the centre."""
        _lib.require_cuda(student_output, teacher_output)
        K = student_output.shape[1]
        B = student_output.shape[0] // self.ncrops
        s = student_output.detach().contiguous()
        t = teacher_output.detach().to(s.dtype).contiguous()
        c = self.center.detach().float().contiguous()
        dev = s.device
        if teacher_output.shape != (2 * B, K) or c.numel() != K or student_output.shape[0] != self.ncrops * B:
            raise ValueError("shape mismatch between student_output, teacher_output and center")
        loss = torch.empty((), dtype=torch.float32, device=dev)
        row_stats = torch.empty((self.ncrops + 2) * B, dtype=torch.float32, device=dev)
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        xchg = self._exchange(K, dev) if world > 1 else None
        # with the peer exchange the kernel writes the column sums straight into the symmetric buffer
        colsum = xchg.de_in.view(-1) if xchg is not None else torch.empty(K, dtype=torch.float32, device=dev)
        grad = torch.empty_like(s)
        g = grad_scale if grad_scale is not None else torch.ones((), dtype=torch.float32, device=dev)
        nbytes = _lib.lib().lafs_dino_fused_workspace_bytes(B, K, self.ncrops)
        if nbytes == 0:
            raise ValueError(f"unsupported DINO shape B={B} K={K} ncrops={self.ncrops}")
        ws = _workspace(dev, nbytes)
        m = float(self.center_momentum)
        temp = float(self.teacher_temp_schedule[epoch])
        main = torch.cuda.current_stream()
        if world == 1:
            new_center = torch.empty(1, K, dtype=torch.float32, device=dev)
            _lib.call("lafs_dino_fwd_bwd", s.data_ptr(), t.data_ptr(), c.data_ptr(), g.data_ptr(), B, K, self.ncrops,
                      1.0 / self.student_temp, 1.0 / temp, _lib.dtype_code(s), loss.data_ptr(), row_stats.data_ptr(),
                      colsum.data_ptr(), grad.data_ptr(), ws.data_ptr(), nbytes, _lib.ptr(new_center),
                      float(np.float32(m)), float(np.float32(1.0 - m)), _lib.stream())
            self.center = new_center
            return loss, grad
        # Several ranks: forward (+ column sums), then the centre exchange -- all-reduce of the [K] column sums
        # (lafs_train.py:675) + centre EMA -- runs on a high-priority SIDE stream while the gradient pass runs on the
        # main stream.  Nothing in the gradient pass needs the new centre (the loss and its gradient use the old one,
        # SURVEY Q7), and the exchange (~25 us) is shorter than the gradient pass (~90 us), so the join after it is free:
        # the step costs what it costs on one GPU.
        # (the fused entry point's workspace is at least as large as the forward-only one)
        _lib.call("lafs_dino_fwd", s.data_ptr(), t.data_ptr(), c.data_ptr(), B, K, self.ncrops,
                  1.0 / self.student_temp, 1.0 / temp, _lib.dtype_code(s), loss.data_ptr(), row_stats.data_ptr(),
                  colsum.data_ptr(), ws.data_ptr(), nbytes, None, 0.0, 0.0, _lib.stream())
        side = getattr(self, "_side", None)
        if side is None or side.device != dev:
            side = self._side = torch.cuda.Stream(device=dev, priority=-1)
        side.wait_stream(main)
        nc = torch.empty(1, K, dtype=torch.float32, device=dev)
        with torch.cuda.stream(side):
            summed = self._allreduce_colsum(colsum)
            _lib.call("lafs_center_ema", c.data_ptr(), summed.data_ptr(), float(2 * B * world), float(np.float32(m)),
                      float(np.float32(1.0 - m)), K, nc.data_ptr(), _lib.stream())
        for tns in (c, colsum, nc, ws):
            tns.record_stream(side)
        _lib.call("lafs_dino_bwd", s.data_ptr(), t.data_ptr(), c.data_ptr(), row_stats.data_ptr(), g.data_ptr(), B, K,
                  self.ncrops, 1.0 / self.student_temp, 1.0 / temp, _lib.dtype_code(s), grad.data_ptr(), _lib.stream())
        main.wait_stream(side)
        self.center = nc
        return loss, grad

    @torch.no_grad()
    def update_center(self, teacher_output):
        """center <- center*m + (all_reduce(sum_rows teacher)/(rows*world))*(1-m)."""
        _lib.require_cuda(teacher_output)
        K = self.center.shape[-1]
        tag = (teacher_output.data_ptr(), teacher_output._version, tuple(teacher_output.shape))
        if self._stash.get("teacher") == tag:
            colsum = self._stash.pop("colsum")       # by-product of forward: no extra pass
            self._stash.pop("teacher", None)
            new_center = self._stash.pop("new_center", None)
            if new_center is not None:               # single process: already computed in the same launch
                self.center = new_center
                return
        else:
            t = teacher_output.detach().contiguous()
            colsum = torch.empty(K, dtype=torch.float32, device=t.device)
            nbytes = 32 * K * 4
            ws = _workspace(t.device, nbytes)
            _lib.call("lafs_colsum", t.data_ptr(), t.shape[0], K, _lib.dtype_code(t), colsum.data_ptr(),
                      ws.data_ptr(), nbytes, _lib.stream())
        world = 1
        if dist.is_available() and dist.is_initialized():
            world = dist.get_world_size()
            if world > 1:
                colsum = self._allreduce_colsum(colsum)
        center = self.center.float().contiguous()
        new_center = torch.empty_like(center)
        m = float(self.center_momentum)
        _lib.call("lafs_center_ema", center.data_ptr(), colsum.data_ptr(),
                  float(len(teacher_output) * world), float(np.float32(m)), float(np.float32(1.0 - m)),
                  K, new_center.data_ptr(), _lib.stream())
        self.center = new_center   # re-bound, like the reference (lafs_train.py:679)
