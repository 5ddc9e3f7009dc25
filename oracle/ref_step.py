"""The hot path executed by the UNMODIFIED reference modules -- TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

Used by bench.py only: `--impl reference` (host cores) and the `reference_eager_b200` report (the same
modules, unpatched, in eager PyTorch on the GPU -- the practical kernel to beat).
The modules come from oracle/ref_harness.py (/root/reference in the build container, the git-ignored
copy baseline/_ref/ on the GPU box).  Every region calls the reference's own callable where one exists:

  extract     VF.extract_patches_pytorch_gridsample (face_pre_pro/ViT_face.py:1615-1656, the
              196-iteration grid_sample loop) + einops rearrange (lafs_train.py:538,544,566)
  embed       nn.Linear(192, dim) = patch_to_embedding (ViT_face.py:619,761), student on every view,
              teacher on the two global views (lafs_train.py:576-579)
  dino        L.DINOLoss.forward (+ update_center) and loss.backward() (lafs_train.py:583,600,643-679)
  ema         the inline loop lafs_train.py:610-613, verbatim
  head        VF.CosFace.forward + nn.CrossEntropyLoss, forward + backward
              (ViT_face.py:49-89, train_largescale.py:604,815-820)

The landmark tail (min-max, +noise, re-sampling) is inline code inside
face_landmark_4simmin_glo_loc.forward (ViT_face.py:1347-1378) behind the out-of-path CNN trunk; it is
restated here line by line (a dozen tiny tensor ops).
"""
import time

import numpy as np
import torch
import torch.nn as nn
from einops import rearrange

from . import ref_harness


class _Timer:
    def __init__(self, device):
        self.cuda = device.type == "cuda"
        self.t = {}

    def run(self, name, fn):
        if self.cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn()
            e1.record()
            e1.synchronize()
            dt = e0.elapsed_time(e1) * 1e-3
        else:
            t0 = time.perf_counter()
            out = fn()
            dt = time.perf_counter() - t0
        self.t[name] = self.t.get(name, 0.0) + dt
        return out


def landmark_tail(raw, n_keep=None):
    """ViT_face.py:1347-1378 with Random_prob=True (+ ran_sample=True when n_keep): same ops, same
    CPU-generator calls (randn on the host, then moved), same order."""
    theta = raw
    t_max = torch.max(theta, dim=1, keepdim=True)[0]
    t_min = torch.min(theta, dim=1, keepdim=True)[0]
    theta = (theta - t_min) / (t_max - t_min) * 111
    theta = theta.view(theta.shape[0], -1, 2)
    theta = theta + (torch.randn(theta.shape) * 5).to(theta.device)
    if n_keep is not None:
        b, c, _ = theta.shape
        extract_id = torch.randint(0, c, (b, n_keep, 1)).repeat(1, 1, 2).to(theta.device)
        theta = torch.gather(theta, 1, extract_id)
    return theta


class SSLReferenceStep:
    """State + one step of the SSL hot path (BASELINE configs[1] shapes by default) on `device`."""

    def __init__(self, device, B, n_local, out_dim, dim, param_shapes, seed=0, amp=None):
        self.ns = ref_harness.load()
        self.dev = torch.device(device)
        self.B, self.L, self.K, self.dim = B, n_local, out_dim, dim
        self.amp = (self.dev.type == "cuda") if amp is None else amp       # reference: fp16 autocast on the GPU
        g = torch.Generator().manual_seed(seed)
        dev = self.dev
        self.img_g = [(torch.rand(B, 3, 112, 112, generator=g) * 2 - 1).to(dev) for _ in range(2)]
        self.img_l = (torch.rand(n_local * B, 3, 112, 112, generator=g) * 2 - 1).to(dev)
        self.raw_g = [torch.randn(B, 392, generator=g).to(dev) for _ in range(2)]
        self.raw_l = torch.randn(n_local * B, 392, generator=g).to(dev)
        ldt = torch.float16 if self.amp else torch.float32
        self.student_out = torch.randn((n_local + 2) * B, out_dim, generator=g).to(dev).to(ldt)
        self.teacher_out = torch.randn(2 * B, out_dim, generator=g).to(dev).to(ldt)
        self.q = [(torch.randn(*s, generator=g) * 0.02).to(dev) for s in param_shapes]
        self.k = [p.clone() for p in self.q]
        self.embed_s = nn.Linear(192, dim).to(dev)
        self.embed_t = nn.Linear(192, dim).to(dev)
        self.loss = self.ns.L.DINOLoss(out_dim, n_local + 2, 0.04, 0.07, 30, 41)
        self.loss = self.loss.cuda() if dev.type == "cuda" else self.loss
        self.patch_shape = torch.tensor([8, 8])

    def step(self, it=0):
        """One pass; returns {region: seconds}."""
        VF = self.ns.VF
        tm = _Timer(self.dev)
        ps = self.patch_shape

        def extract():
            toks = []
            for v in range(2):                                           # lafs_train.py:535-544
                theta = landmark_tail(self.raw_g[v])
                mosaic = VF.extract_patches_pytorch_gridsample(self.img_g[v], theta, ps, num_landm=196)
                toks.append(rearrange(mosaic, 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)', p1=8, p2=8))
            theta = landmark_tail(self.raw_l, 36)                        # lafs_train.py:565-566
            mosaic = VF.extract_patches_pytorch_gridsample(self.img_l, theta, ps, num_landm=36)
            toks.append(rearrange(mosaic, 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)', p1=8, p2=8))
            return toks

        def embed(toks):
            with torch.no_grad(), torch.autocast(self.dev.type, dtype=torch.float16, enabled=self.amp):
                out = [self.embed_s(t) for t in toks]
                out += [self.embed_t(t) for t in toks[:2]]
            return out

        def dino():
            s = self.student_out.detach().requires_grad_(True)
            with torch.autocast(self.dev.type, dtype=torch.float16, enabled=self.amp):
                loss = self.loss(s, self.teacher_out, it % 41)
            loss.backward()
            return loss

        def ema():
            m = 0.996
            with torch.no_grad():                                        # lafs_train.py:610-613, verbatim
                for param_q, param_k in zip(self.q, self.k):
                    param_k.data.mul_(m).add_((1 - m) * param_q.detach().data)

        with torch.no_grad():
            toks = tm.run("extract", extract)
        tm.run("embed", lambda: embed(toks))
        loss = tm.run("dino", dino)
        tm.run("ema", ema)
        tm.t["loss"] = float(loss)
        return tm.t


def head_reference_step(device, B, C, D, iters=3, seed=7):
    """VF.CosFace + CrossEntropyLoss forward+backward (the reference builds the one-hot on the CPU and
    copies it, ViT_face.py:67-82).  Returns seconds per step (median of `iters` after one warm-up)."""
    ns = ref_harness.load()
    dev = torch.device(device)
    torch.manual_seed(seed)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        head = ns.VF.CosFace(D, C, [0] if dev.type == "cuda" else None)
    head = head.to(dev)
    x = torch.randn(B, D, device=dev, requires_grad=True)
    lab = torch.randint(0, C, (B,), device=dev)
    ce = nn.CrossEntropyLoss()
    ts = []
    for i in range(iters + 1):
        x.grad = None
        head.weight.grad = None
        if dev.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss = ce(head(x, lab), lab)
        loss.backward()
        if dev.type == "cuda":
            torch.cuda.synchronize()
        if i > 0:
            ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), float(loss)


def student_update_reference_step(device, param_shapes, iters=3, seed=3):
    """The reference's student update on `device`, eager: utils.clip_gradients (utils.py:132-141, one .item() per
    tensor), torch.optim.AdamW over utils.get_params_groups' two groups (lafs_train.py:396-400), the teacher EMA loop
    (lafs_train.py:610-613).  Returns seconds per step (median of `iters` after one warm-up)."""
    ns = ref_harness.load()
    dev = torch.device(device)
    g = torch.Generator().manual_seed(seed)

    class Holder(nn.Module):
        def __init__(self):
            super().__init__()
            self.ps = nn.ParameterList([nn.Parameter((torch.randn(*s, generator=g) * 0.02).to(dev)) for s in param_shapes])

    student = Holder()
    teacher = [p.detach().clone() for p in student.ps]
    reg = [p for p in student.ps if p.dim() > 1]
    noreg = [p for p in student.ps if p.dim() <= 1]
    opt = torch.optim.AdamW([{"params": reg}, {"params": noreg, "weight_decay": 0.}])
    grads = [(torch.randn(*s, generator=g) * 0.01).to(dev) for s in param_shapes]
    ts = []
    for i in range(iters + 1):
        for p, gr in zip(student.ps, grads):
            p.grad = gr.clone()
        if dev.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        ns.dutils.clip_gradients(student, 3.0)
        opt.step()
        with torch.no_grad():
            m = 0.996
            for param_q, param_k in zip(student.ps, teacher):
                param_k.data.mul_(m).add_((1 - m) * param_q.detach().data)
        if dev.type == "cuda":
            torch.cuda.synchronize()
        if i > 0:
            ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def dino_head_reference_step(device, B, ncrops, K, D, iters=2, seed=11):
    """The reference's DINOHead tail + DINOLoss from the bottleneck features, forward + backward, on `device`:
    x = F.normalize(x); x = last_layer(x) (vision_transformer.py:298-300, weight-normed Linear built by the reference's
    own DINOHead constructor) for the student (all crops) and the teacher (2 global crops), L.DINOLoss.forward
    (+ update_center) and loss.backward() (lafs_train.py:583,600,643-679).  Returns (seconds per step, loss)."""
    import warnings
    ns = ref_harness.load()
    dev = torch.device(device)
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        hs = ns.vt.DINOHead(D, K, nlayers=1, bottleneck_dim=D).to(dev)      # mlp = one Linear (not timed), last_layer D -> K
        ht = ns.vt.DINOHead(D, K, nlayers=1, bottleneck_dim=D).to(dev)
    crit = ns.L.DINOLoss(K, ncrops, 0.04, 0.04, 0, 1).to(dev)
    xs = torch.randn(ncrops * B, D, device=dev, requires_grad=True)
    xt = torch.randn(2 * B, D, device=dev)
    ts = []
    for i in range(iters + 1):
        xs.grad = None
        hs.zero_grad(set_to_none=True)
        if dev.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.no_grad():
            t_out = ht.last_layer(nn.functional.normalize(xt, dim=-1, p=2))
        s_out = hs.last_layer(nn.functional.normalize(xs, dim=-1, p=2))
        loss = crit(s_out, t_out, 0)
        loss.backward()
        if dev.type == "cuda":
            torch.cuda.synchronize()
        if i > 0:
            ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), float(loss.detach())
