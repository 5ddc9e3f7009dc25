"""Generates tests/golden/dino_head.npz by running the UNMODIFIED reference (/root/reference) on CPU:
vision_transformer.DINOHead (student and teacher) + lafs_train.DINOLoss, the loss of lafs_train.py:581-583 and its
gradients w.r.t. the bottleneck features and last_layer.weight_v.  Run once in the build container:
    python tests/golden/make_golden_dino_head.py
(kept apart from make_golden.py so that the other fixtures are not rewritten)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    ns = ref_harness.load()
    vt, L = ns.vt, ns.L
    torch.set_num_threads(1)
    torch.manual_seed(21)
    g = torch.Generator().manual_seed(4321)
    in_dim, hidden, D, K, B, ncrops = 48, 96, 64, 1200, 4, 6
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        student = vt.DINOHead(in_dim, K, nlayers=3, hidden_dim=hidden, bottleneck_dim=D, norm_last_layer=False)
        teacher = vt.DINOHead(in_dim, K, nlayers=3, hidden_dim=hidden, bottleneck_dim=D)
    # random (not unit) prototype scales so that weight_g matters; spread the prototypes
    with torch.no_grad():
        for h in (student, teacher):
            h.last_layer.weight_v.copy_(torch.randn(K, D, generator=g) * 0.5)
        student.last_layer.weight_g.copy_(0.75 + 0.5 * torch.rand(K, 1, generator=g))
        teacher.last_layer.weight_g.copy_(0.75 + 0.5 * torch.rand(K, 1, generator=g))
    box = {}

    def keep_student(m, i, o):
        o.retain_grad()
        box["xs"] = o

    def keep_teacher(m, i, o):
        box["xt"] = o.detach().clone()

    student.mlp.register_forward_hook(keep_student)
    teacher.mlp.register_forward_hook(keep_teacher)
    dl = L.DINOLoss(K, ncrops, 0.04, 0.07, 30, 41)
    dl.center = torch.randn(1, K, generator=g) * 0.05
    center0 = dl.center.clone()
    feat_s = torch.randn(ncrops * B, in_dim, generator=g) * 3
    feat_t = torch.randn(2 * B, in_dim, generator=g) * 3
    epoch = 5
    with torch.no_grad():
        t_out = teacher(feat_t)
    s_out = student(feat_s)
    loss = dl(s_out, t_out, epoch)
    loss.backward()
    out = dict(xs=box["xs"].detach(), xt=box["xt"], vs=student.last_layer.weight_v.detach(),
               gs=student.last_layer.weight_g.detach(), vt=teacher.last_layer.weight_v.detach(),
               gt=teacher.last_layer.weight_g.detach(), center0=center0, center1=dl.center, loss=loss.detach(),
               grad_xs=box["xs"].grad, grad_vs=student.last_layer.weight_v.grad, grad_gs=student.last_layer.weight_g.grad,
               student_logits_row0=s_out[0].detach(), teacher_logits_row0=t_out[0], epoch=epoch,
               temp=dl.teacher_temp_schedule[epoch], ncrops=ncrops)
    out = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()}
    np.savez_compressed(os.path.join(OUT, "dino_head.npz"), **out)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items()}, float(out["loss"]))


if __name__ == "__main__":
    main()
