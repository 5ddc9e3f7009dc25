"""Teacher EMA as ONE multi-tensor launch.

Mirrors the inline loop of the reference training step (lafs_train.py:610-613):

    for param_q, param_k in zip(student.module.parameters(), teacher_without_ddp.parameters()):
        param_k.data.mul_(m).add_((1 - m) * param_q.detach().data)

`ema_update_(teacher_params, student_params, m)` is semantically (and bit-for-bit) the same,
in place, under no_grad.  The pointer table is cached per parameter list.
"""
import numpy as np
import torch

from . import _lib

_REC = np.dtype([("k", "<u8"), ("q", "<u8"), ("n", "<i8")])


class EmaPlan:
    """Device-resident chunk table for a fixed list of (teacher, student) fp32 tensors."""

    def __init__(self, teacher_params, student_params):
        ks, qs = list(teacher_params), list(student_params)
        if len(ks) != len(qs):
            # the reference's zip() silently truncates; a mismatch is a bug worth surfacing
            raise ValueError(f"teacher has {len(ks)} tensors, student {len(qs)}")
        recs = []
        self.numel = 0
        for k, q in zip(ks, qs):
            k, q = k.data, q.data
            _lib.require_cuda(k, q)
            if k.dtype != torch.float32 or q.dtype != torch.float32:
                raise TypeError("ema_update_: fp32 parameters expected (the reference keeps fp32 masters)")
            if k.shape != q.shape or not k.is_contiguous() or not q.is_contiguous():
                raise ValueError("ema_update_: parameter pairs must be contiguous and equally shaped")
            n, kp, qp = k.numel(), k.data_ptr(), q.data_ptr()
            self.numel += n
            for off in range(0, n, _lib.EMA_CHUNK):
                recs.append((kp + 4 * off, qp + 4 * off, min(_lib.EMA_CHUNK, n - off)))
        self.nchunks = len(recs)
        self._keep = (ks, qs)            # the table holds raw pointers: keep the tensors alive
        self.key = self.make_key(ks, qs)
        dev = ks[0].device if ks else torch.device("cuda")
        table = np.array(recs, dtype=_REC)
        self.table = torch.from_numpy(table.view(np.uint8).copy()).to(dev) if recs else None

    @staticmethod
    def make_key(ks, qs):
        return tuple((k.data_ptr(), q.data_ptr(), k.numel()) for k, q in zip(ks, qs))

    def step(self, m, max_ctas=0):
        """max_ctas > 0: persistent launch with that many CTAs (co-scheduling next to another kernel)."""
        if self.nchunks == 0:
            return
        m = float(m)
        # PyTorch rounds the python/numpy float64 scalars m and (1-m) to fp32 separately
        _lib.call("lafs_ema_multi", self.table.data_ptr(), self.nchunks,
                  float(np.float32(m)), float(np.float32(1.0 - m)), int(max_ctas), _lib.stream())


_plans = {}


@torch.no_grad()
def ema_update_(teacher_params, student_params, m):
    """In-place k <- m*k + (1-m)*q for every pair.  Accepts parameters or tensors."""
    ks, qs = list(teacher_params), list(student_params)
    key = EmaPlan.make_key([k.data for k in ks], [q.data for q in qs])
    plan = _plans.get(key)
    if plan is None:
        if len(_plans) > 16:
            _plans.clear()
        plan = _plans[key] = EmaPlan(ks, qs)
    plan.step(m)
    return plan
