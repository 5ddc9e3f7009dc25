"""Student update of the SSL step as ONE multi-tensor call: per-tensor gradient clipping, AdamW and the teacher
EMA.  Mirrors, in this order, what lafs_train.py:601-613 does per step:

    param_norms = utils.clip_gradients(student, args.clip_grad)          # utils.py:132-141, one host sync per tensor
    utils.cancel_gradients_last_layer(epoch, student, freeze_last_layer) # utils.py:144-149 -> grads[i] = None here
    optimizer.step()                                                     # torch.optim.AdamW, 2 param groups (utils.py:662-673)
    for param_q, param_k in zip(student.parameters(), teacher.parameters()):
        param_k.data.mul_(m).add_((1 - m) * param_q.detach().data)       # lafs_train.py:610-613

`StudentUpdate.step(grads, lr, weight_decay, clip_grad, ema_momentum)` does all of it in two passes over the
parameters (csrc/optim.cu) and returns the per-tensor gradient norms as a DEVICE tensor (the reference's list of
Python floats costs 147 `.item()` syncs per step).  Moments (`exp_avg`, `exp_avg_sq`) and the step counters follow
torch.optim.AdamW (a tensor without a gradient in a step is skipped entirely, like AdamW skips `p.grad is None`).
"""
import numpy as np
import torch

from . import _lib

_REC = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("k", "<u8"), ("n", "<i4"), ("t", "<i4")])


def regularized_mask(named_parameters):
    """utils.get_params_groups (utils.py:662-673): biases and 1-D (norm) parameters are not weight-decayed."""
    return [not (name.endswith(".bias") or p.dim() == 1) for name, p in named_parameters]


class StudentUpdate:
    def __init__(self, student_params, teacher_params=None, regularized=None, betas=(0.9, 0.999), eps=1e-8):
        self.params = [p.data if isinstance(p, torch.nn.Parameter) else p for p in student_params]
        self.teacher = None if teacher_params is None else [k.data if isinstance(k, torch.nn.Parameter) else k
                                                            for k in teacher_params]
        if self.teacher is not None and len(self.teacher) != len(self.params):
            raise ValueError(f"teacher has {len(self.teacher)} tensors, student {len(self.params)}")
        for t in self.params + (self.teacher or []):
            _lib.require_cuda(t)
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise TypeError("StudentUpdate: contiguous fp32 parameters expected (the reference keeps fp32 masters)")
        n = len(self.params)
        self.regularized = [True] * n if regularized is None else [bool(r) for r in regularized]
        if len(self.regularized) != n:
            raise ValueError("regularized mask must have one entry per parameter tensor")
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self.steps = [0] * n                       # per tensor, like AdamW's state['step']
        dev = self.params[0].device if n else torch.device("cuda")
        self.dev = dev
        self.grad_norms = torch.zeros(max(n, 1), dtype=torch.float32, device=dev)
        self.clip_coef = torch.ones(max(n, 1), dtype=torch.float32, device=dev)
        self.reg_dev = torch.tensor(self.regularized, dtype=torch.uint8, device=dev) if n else None
        self.hyper = torch.zeros(10, dtype=torch.float32, device=dev)
        self._hyper_host = torch.zeros(10, dtype=torch.float32).pin_memory()
        self._key = None
        self._table = self._first = self._ws = None
        self.nchunks = 0

    def _build(self, grads):
        recs, first = [], [0]
        for t, p in enumerate(self.params):
            g = grads[t]
            if g is not None:
                if g.shape != p.shape or g.dtype != torch.float32 or not g.is_contiguous() or g.device != p.device:
                    raise ValueError(f"gradient {t}: contiguous fp32 tensor of shape {tuple(p.shape)} on {p.device} expected")
            k = None if self.teacher is None else self.teacher[t]
            n = p.numel()
            for off in range(0, n, _lib.EMA_CHUNK):
                b = 4 * off
                recs.append((p.data_ptr() + b, 0 if g is None else g.data_ptr() + b, self.exp_avg[t].data_ptr() + b,
                             self.exp_avg_sq[t].data_ptr() + b, 0 if k is None else k.data_ptr() + b,
                             min(_lib.EMA_CHUNK, n - off), t))
            first.append(len(recs))
        self.nchunks = len(recs)
        table = np.array(recs, dtype=_REC)
        self._table = torch.from_numpy(table.view(np.uint8).copy()).to(self.dev)
        self._first = torch.tensor(first, dtype=torch.int32, device=self.dev)
        nbytes = _lib.lib().lafs_optim_workspace_bytes(self.nchunks, len(self.params))
        self._ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=self.dev)
        self._keep = list(grads)

    @torch.no_grad()
    def step(self, grads, lr, weight_decay, clip_grad=0.0, ema_momentum=None):
        """grads: one fp32 tensor (or None = no gradient this step) per parameter tensor.  clip_grad <= 0 / None:
        no clipping.  ema_momentum None: the teacher is left alone.  Returns the device tensor of per-tensor
        gradient norms (0 for tensors without a gradient)."""
        if not self.params:
            return self.grad_norms[:0]
        grads = [None if g is None else g.detach() for g in grads]
        if len(grads) != len(self.params):
            raise ValueError("one gradient entry per parameter tensor expected")
        if self.teacher is None and ema_momentum is not None:
            raise ValueError("ema_momentum given but no teacher parameters")
        key = tuple(0 if g is None else g.data_ptr() for g in grads)
        if key != self._key:
            self._build(grads)
            self._key = key
        # AdamW keeps one step counter per tensor; tensors that always have (or never have) a gradient share it.
        # The kernel takes ONE (step_size, bc2_sqrt) pair: the counters of the tensors updated in this call must agree.
        live = [t for t, g in enumerate(grads) if g is not None]
        for t in live:
            self.steps[t] += 1
        counts = {self.steps[t] for t in live}
        if len(counts) > 1:
            raise ValueError("tensors updated in one call have different AdamW step counts; split the call "
                             "(e.g. after un-freezing the last layer, utils.py:144-149) or call reset_steps()")
        step = counts.pop() if counts else 1
        b1, b2 = self.betas
        lr, wd = float(lr), float(weight_decay)
        m = 1.0 if ema_momentum is None else float(ema_momentum)
        h = self._hyper_host
        h[0] = 1.0 - lr * wd
        h[1] = lr / (1.0 - b1 ** step)
        h[2] = (1.0 - b2 ** step) ** 0.5
        h[3] = self.eps
        h[4] = 1.0 - b1
        h[5] = b2
        h[6] = 1.0 - b2
        h[7] = float(np.float32(m))
        h[8] = float(np.float32(1.0 - m))
        h[9] = float(clip_grad) if clip_grad else 0.0
        self.hyper.copy_(h, non_blocking=True)
        _lib.call("lafs_adamw_ema_multi", self._table.data_ptr(), self.nchunks, self._first.data_ptr(), self.reg_dev.data_ptr(),
                  len(self.params), self.hyper.data_ptr(), self.grad_norms.data_ptr(), self.clip_coef.data_ptr(),
                  self._ws.data_ptr(), self._ws.numel(), _lib.stream())
        return self.grad_norms
