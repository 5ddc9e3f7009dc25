"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run once in the build container:  python tests/golden/make_golden.py
The reference cannot travel to the GPU box, so its outputs on small seeded inputs are
committed here; tests compare both the oracle (oracle/lafs_oracle.py) and the CUDA path
against them.  Every array is produced by reference code, cited per case.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: (v.shape, str(v.dtype)) for k, v in out.items()})


def main():
    ns = ref_harness.load()
    VF, L, dutils, mixup = ns.VF, ns.L, ns.dutils, ns.mixup
    torch.set_num_threads(1)  # fixtures independent of threading

    # ---- (1a) extract_patches_pytorch_gridsample, ViT_face.py:1615-1656 ------------------
    g = torch.Generator().manual_seed(1234)
    imgs = torch.rand(2, 3, 112, 112, generator=g) * 2 - 1
    th196 = torch.rand(2, 196, 2, generator=g) * 111 + torch.randn(2, 196, 2, generator=g) * 5
    # force border / fully-outside / exact-integer / half-integer landmarks
    th196[0, 0] = torch.tensor([0.0, 0.0]); th196[0, 1] = torch.tensor([111.0, 111.0])
    th196[0, 2] = torch.tensor([-20.0, 50.0]); th196[0, 3] = torch.tensor([130.5, -3.25])
    th196[0, 4] = torch.tensor([4.5, 4.5]); th196[0, 5] = torch.tensor([56.0, 55.5])
    th196[1, 0] = torch.tensor([3.999999, 107.50001]); th196[1, 1] = torch.tensor([-4.5, 115.5])
    th36 = th196[:, torch.randint(0, 196, (36,), generator=g)].contiguous()
    ps = torch.tensor([8, 8])
    out196 = VF.extract_patches_pytorch_gridsample(imgs, th196, ps, 196)
    out36 = VF.extract_patches_pytorch_gridsample(imgs, th36, ps, 36)
    save("patches", imgs=imgs, theta196=th196, theta36=th36, mosaic196=out196, mosaic36=out36)

    # ---- (1b) landmark CNN tail: min-max, noise, gather.  ViT_face.py:1316-1409 -----------
    torch.manual_seed(7)
    with contextlib.redirect_stdout(io.StringIO()):
        cnn = VF.face_landmark_4simmin_glo_loc(
            loss_type="CosFace", GPU_ID=None, num_class=10, num_patches=196, image_size=112,
            patch_size=8, dim=512, depth=1, heads=2, mlp_dim=64, dropout=0.0, emb_dropout=0.0)
    cnn.eval()
    raw_box = {}
    cnn.output_layer.register_forward_hook(lambda m, i, o: raw_box.__setitem__("raw", o.detach().clone()))
    x = torch.rand(3, 3, 112, 112, generator=g) * 2 - 1
    x_aug = torch.rand(3, 3, 112, 112, generator=g) * 2 - 1
    with torch.no_grad():
        torch.manual_seed(100)   # global view: noise only (lafs_train.py:535)
        th_g, mos_g = cnn(x, x_Aug=x_aug, patch_shape=ps, Random_prob=True, return_prob=True)
        raw_g = raw_box["raw"]
        torch.manual_seed(101)   # local view: noise + 36 re-sampled landmarks (lafs_train.py:565)
        th_l, mos_l = cnn(x, x_Aug=x_aug, patch_shape=ps, Random_prob=True, ran_sample=True)
        raw_l = raw_box["raw"]
        th_p, mos_p = cnn(x, patch_shape=ps)   # plain (no noise): finetune-like tail
    save("landmark_post", raw_global=raw_g, theta_global=th_g, seed_global=100,
         raw_local=raw_l, theta_local=th_l, seed_local=101, theta_plain=th_p,
         x_aug=x_aug.half(),  # fp16-exact images keep the file small
         mosaic_local_from_half=VF.extract_patches_pytorch_gridsample(x_aug.half().float(), th_l, ps, 36))

    # ---- (1c) patch_to_embedding on tokens, ViT_face.py:759-761 + lafs_train.py:538 --------
    from einops import rearrange
    torch.manual_seed(8)
    with contextlib.redirect_stdout(io.StringIO()):
        vit = VF.ViT_face_landmark_patch8(
            loss_type="CosFace", GPU_ID=None, num_class=16, num_patches=196, image_size=112,
            patch_size=8, dim=64, depth=1, heads=2, mlp_dim=64, dropout=0.0, emb_dropout=0.0,
            with_land=False)
    vit.eval()
    emb_box = {}
    vit.patch_to_embedding.register_forward_hook(lambda m, i, o: emb_box.__setitem__("y", o.detach().clone()))
    tok = rearrange(out196, "b c (h p1) (w p2) -> b (h w) (p1 p2 c)", p1=8, p2=8)
    with torch.no_grad():
        vit(tok)
    save("patch_embed", weight=vit.patch_to_embedding.weight, bias=vit.patch_to_embedding.bias,
         tokens_checksum=tok.double().sum(), tokens_row0=tok[0, :4], embedded=emb_box["y"])

    # ---- (2) DINOLoss, lafs_train.py:626-679 ---------------------------------------------
    B, K, ncrops = 4, 1024, 6
    dl = L.DINOLoss(K, ncrops, 0.04, 0.07, 30, 41)
    dl.center = torch.randn(1, K, generator=g) * 0.1
    center0 = dl.center.clone()
    s = (torch.randn(ncrops * B, K, generator=g) * 1.5).requires_grad_(True)
    t = torch.randn(2 * B, K, generator=g) * 1.5
    epoch = 3
    loss = dl(s, t, epoch)
    loss.backward()
    save("dino", student=s, teacher=t, center0=center0, center1=dl.center, loss=loss,
         grad_student=s.grad, epoch=epoch, temp=dl.teacher_temp_schedule[epoch],
         temp_schedule=dl.teacher_temp_schedule, ncrops=ncrops)
    # bf16-valued inputs (exactly representable): the case the CUDA bf16 path is compared on
    sb = (torch.randn(ncrops * B, K, generator=g) * 2).bfloat16().float().requires_grad_(True)
    tb = (torch.randn(2 * B, K, generator=g) * 2).bfloat16().float()
    dl2 = L.DINOLoss(K, ncrops, 0.04, 0.07, 30, 41)
    dl2.center = center0.clone()
    loss2 = dl2(sb, tb, 35)
    loss2.backward()
    save("dino_bf16", student=sb.detach().bfloat16().view(torch.int16), teacher=tb.bfloat16().view(torch.int16),
         center0=center0, center1=dl2.center, loss=loss2, grad_student=sb.grad, epoch=35,
         temp=dl2.teacher_temp_schedule[35], ncrops=ncrops)

    # ---- (3) teacher EMA, lafs_train.py:610-613 + utils.py:187-198 ------------------------
    sched = dutils.cosine_scheduler(0.996, 1, 41, 100)
    shapes = [(256,), (1, 197, 64), (64, 192), (264, 64), (7,), (1, 1, 3), (129, 33)]
    q = [torch.randn(*sh, generator=g) for sh in shapes]
    k0 = [torch.randn(*sh, generator=g) for sh in shapes]
    k = [a.clone() for a in k0]
    it = 17
    m = sched[it]
    with torch.no_grad():
        for param_q, param_k in zip(q, k):
            param_k.data.mul_(m).add_((1 - m) * param_q.detach().data)
    save("ema", m=np.float64(m), it=it, sched_head=sched[:64],
         **{f"q{i}": a for i, a in enumerate(q)}, **{f"k0_{i}": a for i, a in enumerate(k0)},
         **{f"k1_{i}": a for i, a in enumerate(k)})

    # ---- (4) CosFace + CE, ViT_face.py:26-96, train_largescale.py:601-604 ------------------
    torch.manual_seed(9)
    Bh, D, C = 8, 64, 1000
    with contextlib.redirect_stdout(io.StringIO()):
        head = VF.CosFace(in_features=D, out_features=C, device_id=None)
    xh = torch.randn(Bh, D, generator=g).requires_grad_(True)
    lab = torch.randint(0, C, (Bh,), generator=g)
    lab[0] = 0; lab[1] = C - 1
    logits = head(xh, lab)
    loss_h = torch.nn.CrossEntropyLoss()(logits, lab)
    loss_h.backward()
    gx_h, gw_h = xh.grad.clone(), head.weight.grad.clone()
    xh.grad = None; head.weight.grad = None
    lam = 0.3
    soft = mixup.mixup_target(lab, C, lam=lam, smoothing=0.0, device="cpu")
    logits_s = head(xh, soft)
    loss_s = torch.sum(-soft * torch.nn.functional.log_softmax(logits_s, dim=-1), dim=-1).mean()
    loss_s.backward()
    save("cosface", x=xh, weight=head.weight, label=lab, logits_hard=logits, loss_hard=loss_h,
         grad_x_hard=gx_h, grad_w_hard=gw_h, lam=lam, soft_target_nnz=(soft != 0).sum(1),
         label_b=lab.flip(0), logits_soft=logits_s, loss_soft=loss_s,
         grad_x_soft=xh.grad, grad_w_soft=head.weight.grad)

    # ---- (4b) torch.chunk shard sizes, ViT_face.py:56 -------------------------------------
    rows = []
    for Cn in (93431, 205990, 1000, 10, 7):
        for R in (1, 2, 3, 4, 8):
            sizes = [c.shape[0] for c in torch.chunk(torch.empty(Cn, 1), R, dim=0)]
            rows.append([Cn, R] + sizes + [0] * (8 - len(sizes)))
    save("shards", table=np.array(rows, dtype=np.int64))


if __name__ == "__main__":
    main()
