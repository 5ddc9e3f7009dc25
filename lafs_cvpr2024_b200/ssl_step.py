"""Host-side composition of the SSL hot path for one training step (what
lafs_train.py:train_one_epoch does between the data loader and the optimizer, minus the
transformer blocks / DINO head / landmark-CNN trunk, which stay stock PyTorch and are outside
the path -- SURVEY.md section 8).

Order of operations per step (reference line numbers in lafs_train.py):
  :535,:541  landmark tail (min-max, +noise) for the two global views      -> theta_g [2B,196,2]
  :565       landmark tail (+noise, 36 re-sampled landmarks) for L locals  -> theta_l [LB,36,2]
  :538-:569  patch gather (+ rearrange) for every view                      -> tokens
  :583       DINOLoss forward (+ centre update), :600 its backward
  :610-613   teacher EMA over every parameter pair
"""
import torch

from . import _lib
from .dino_loss import DINOLoss
from .ema import EmaPlan
from .patches import (PatchEmbedWeights, embed_backward_weight, extract_tokens, gather_embed, landmark_post,
                      new_token_buffer)


class SSLHotPath:
    """Holds the persistent state of the path (centre, EMA plan) and runs one step on the
    current CUDA stream.  All inputs are CUDA tensors; nothing here touches the host."""

    def __init__(self, out_dim, n_local, teacher_params, student_params, nepochs=41,
                 warmup_teacher_temp=0.04, teacher_temp=0.07, warmup_teacher_temp_epochs=30,
                 student_embed=None, teacher_embed=None):
        """student_embed / teacher_embed: (weight [dim,192], bias [dim]) of the two
        patch_to_embedding layers (ViT_face.py:619); enables the fused gather->embed path."""
        self.n_local = n_local
        self.embed_global = self.embed_local = None
        if student_embed is not None:
            self.embed_global = PatchEmbedWeights([student_embed, teacher_embed])   # 2 models, 1 gather
            self.embed_local = PatchEmbedWeights([student_embed])
        self.loss = DINOLoss(out_dim, n_local + 2, warmup_teacher_temp, teacher_temp,
                             warmup_teacher_temp_epochs, nepochs).cuda()
        self.ema = EmaPlan(list(teacher_params), list(student_params))

    def landmarks_and_tokens(self, raw_g, noise_g, img_g, raw_l, noise_l, idx_l, img_l):
        """raw_* [*,392] landmark regressions, noise_* [*,196,2], idx_l [LB,36] int64,
        img_* the augmented views the patches are cut from."""
        theta_g = landmark_post(raw_g, noise_g)
        tok_g = extract_tokens(img_g, theta_g)
        theta_l = landmark_post(raw_l, noise_l, idx_l)
        tok_l = extract_tokens(img_l, theta_l)
        return theta_g, tok_g, theta_l, tok_l

    def landmarks_and_embeddings(self, raw_g, noise_g, img_g, raw_l, noise_l, idx_l, img_l, refresh=True, keep_tokens=False):
        """Fused form: landmark tail, then gather -> patch_to_embedding on the tensor cores.
        Returns (student_global, teacher_global, student_local) embedded tokens (bf16); the
        patch mosaics / fp32 token tensors of the reference are never materialised.
        keep_tokens: also keep the gathered bf16 tokens (persistent buffers) for student_embed_backward."""
        if refresh:
            self.embed_global.refresh()    # weights moved by the optimizer / EMA since last step
            self.embed_local.refresh()
        tg = tl = None
        if keep_tokens:
            ng, nl = noise_g.shape[0] * noise_g.shape[1], idx_l.shape[0] * idx_l.shape[1]
            if getattr(self, "_tok_g", None) is None or self._tok_g.shape[0] != ng or self._tok_l.shape[0] != nl:
                self._tok_g, self._tok_l = new_token_buffer(ng, raw_g.device), new_token_buffer(nl, raw_g.device)
            tg, tl = self._tok_g, self._tok_l
        theta_g = landmark_post(raw_g, noise_g)
        s_g, t_g = gather_embed(img_g, theta_g, self.embed_global, save_tokens=tg)
        theta_l = landmark_post(raw_l, noise_l, idx_l)
        (s_l,) = gather_embed(img_l, theta_l, self.embed_local, save_tokens=tl)
        return s_g, t_g, s_l

    def student_embed_backward(self, grad_s_g, grad_s_l):
        """Weight / bias gradient of the STUDENT's patch_to_embedding (lafs_train.py:600 through ViT_face.py:761)
        from the gradients of its embedded tokens (global [2B,196,dim], local [LB,36,dim], bf16) and the tokens kept
        by landmarks_and_embeddings(keep_tokens=True): two split-K tcgen05 GEMMs accumulating into one
        [dim,192] + [dim] result.  (The landmark CNN is frozen in the SSL stage, lafs_train.py:146-150: there is no
        gradient w.r.t. the landmarks or the images.)"""
        gw, gb = embed_backward_weight(grad_s_g, self._tok_g)
        return embed_backward_weight(grad_s_l, self._tok_l, grad_w=gw, grad_b=gb, accumulate=True)

    def loss_and_grad(self, student_out, teacher_out, epoch, fused=True):
        """DINO loss, its gradient w.r.t. the student logits, and the centre update.  fused=True uses
        the single-call forward+backward entry point; fused=False goes through autograd."""
        if fused:
            return self.loss.loss_and_grad(student_out, teacher_out, epoch)
        s = student_out.detach().requires_grad_(True)
        loss = self.loss(s, teacher_out, epoch)
        loss.backward()
        return loss.detach(), s.grad

    def ema_step(self, m, max_ctas=0):
        self.ema.step(m, max_ctas)


class GraphedSSLStep:
    """The whole hot-path step captured once into a CUDA graph and replayed with a single launch
    (11 kernels -> 1 graph launch; the inputs live in static device buffers that the caller
    refreshes, e.g. with copy_ from pinned host memory on a copy stream).

    Static inputs : img_g, img_l (uint8 or fp32), raw_g, raw_l, noise_g, noise_l, idx_l,
                    student_out, teacher_out, optionally grad_s_g / grad_s_l (gradients of the student's embedded
                    tokens: adds the patch_to_embedding weight/bias gradient to the step).
    Outputs: loss (0-dim), grad_student, the three embedded-token tensors, grad_embed_w / grad_embed_b.  The DINO centre is kept in a static buffer and updated in place at the
    end of the graph (the loss and its backward inside the graph still see the old centre, Q7)."""

    def __init__(self, path: SSLHotPath, static_inputs: dict, epoch: int, momentum: float, overlap_ema=False, ema_ctas=444,
                 fused_loss=True, center=None):
        """overlap_ema: put the teacher EMA (as `ema_ctas` persistent CTAs; 0 = one CTA per chunk) on a second
        captured stream.  True / "early": forked before the gather->embed kernels (measured on B200 in round 1:
        between -3 % and +17 % step time; the gather CTA fills an SM's registers, so the EMA cannot co-reside with it).
        "late": forked AFTER the gather->embed kernels, next to the patch-embed backward GEMMs and the DINO kernels,
        which leave 45-55 % of the HBM bandwidth idle -- the EMA (HBM-bound at 0.95) fills it."""
        self.path, self.inp = path, static_inputs
        self.overlap_ema = overlap_ema
        self.fused_loss = fused_loss
        self.ema_ctas = ema_ctas
        self.side = torch.cuda.Stream()
        # `center`: an existing static centre buffer to share (several graphs over different input buffers --
        # e.g. the two halves of a double-buffered loader hand-off -- that advance ONE training state)
        self.center = center if center is not None else path.loss.center.detach().clone().contiguous()
        self.graph = torch.cuda.CUDAGraph()
        self.out = {}
        # Warm-up outside capture (lazy init, workspaces) must not change training state: it runs with
        # momentum 1.0 (k*1 + q*0 leaves the teacher bit-identical for finite q) and the centre is put back
        # afterwards, so constructing the graph is side-effect free whatever the static input buffers hold.
        # Capture itself executes nothing.  The static inputs must hold valid data before the first replay().
        center0 = self.center.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self._body(epoch, 1.0)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.center.copy_(center0)
        with torch.cuda.graph(self.graph):
            self._body(epoch, momentum)

    def _body(self, epoch, momentum):
        p, i = self.path, self.inp
        p.loss.center = self.center
        main = torch.cuda.current_stream()
        early = self.overlap_ema in (True, "early")
        late = self.overlap_ema in ("late", "late2")
        side2 = self.overlap_ema == "late2"      # "late2": the patch-embed backward GEMMs get their own stream as well
        if early:
            # the EMA reads the teacher's patch_to_embedding weights that the weight-prep kernels
            # also read, and writes them: run the prep first, then fork
            p.embed_global.refresh()
            p.embed_local.refresh()
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                p.ema_step(momentum, max_ctas=self.ema_ctas)
        with_bwd = "grad_s_g" in i          # synthetic gradients of the student's embedded tokens (static inputs)
        s_g, t_g, s_l = p.landmarks_and_embeddings(i["raw_g"], i["noise_g"], i["img_g"], i["raw_l"], i["noise_l"],
                                                   i["idx_l"], i["img_l"], refresh=not early, keep_tokens=with_bwd)
        if late:      # the teacher's embedding weights have been read by the kernels above: the EMA may now move them
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                p.ema_step(momentum, max_ctas=self.ema_ctas)
        gw = gb = None
        if with_bwd and side2:
            if getattr(self, "side_b", None) is None:
                self.side_b = torch.cuda.Stream()
            self.side_b.wait_stream(main)
            with torch.cuda.stream(self.side_b):
                gw, gb = p.student_embed_backward(i["grad_s_g"], i["grad_s_l"])
        elif with_bwd:                       # student patch_to_embedding backward (lafs_train.py:600 / ViT_face.py:761)
            gw, gb = p.student_embed_backward(i["grad_s_g"], i["grad_s_l"])
        loss, grad = p.loss_and_grad(i["student_out"], i["teacher_out"], epoch, fused=self.fused_loss)
        self.center.copy_(p.loss.center)              # static centre buffer <- re-bound new centre
        p.loss.center = self.center
        if self.overlap_ema:
            main.wait_stream(self.side)
            if side2 and with_bwd:
                main.wait_stream(self.side_b)
        else:
            p.ema_step(momentum)
        self.out = {"loss": loss.detach(), "grad_student": grad, "s_g": s_g, "t_g": t_g, "s_l": s_l,
                    "grad_embed_w": gw, "grad_embed_b": gb}

    def replay(self):
        self.graph.replay()
        return self.out["loss"]
