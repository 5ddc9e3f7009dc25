"""Landmark tail and bilinear patch gather with the reference's function surface.

    extract_patches_pytorch_gridsample(imgs, landmarks, patch_shape, num_landm=49)
        face_pre_pro/ViT_face.py:1615-1656 -- returns the [B,C,8r,8r] mosaic, differentiable
        w.r.t. imgs and landmarks (the finetune path back-propagates into the landmark CNN).
    extract_tokens(imgs, landmarks)            mosaic + einops rearrange of lafs_train.py:538 fused
    landmark_post(raw, noise=None, extract_id=None)   ViT_face.py:1347-1378
"""
import torch

from . import _lib

# LAFS_COORD_DIV reproduces the reference's CPU arithmetic (and the committed golden vectors)
# bit for bit; set to _lib.COORD_RECIP to reproduce eager-CUDA's reciprocal multiply instead.
COORD_MODE = _lib.COORD_DIV


class _GatherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, imgs, theta, layout, coord_mode):
        _lib.require_cuda(imgs, theta)
        if imgs.dim() != 4 or theta.dim() != 3 or theta.shape[-1] != 2 or theta.shape[0] != imgs.shape[0]:
            raise ValueError(f"imgs [B,C,H,W] / landmarks [B,n,2] expected, got {tuple(imgs.shape)} {tuple(theta.shape)}")
        x = imgs.detach().float().contiguous()
        th = theta.detach().float().contiguous()
        Bv, Cc, H, W = x.shape
        n = th.shape[1]
        if layout == _lib.LAYOUT_TOKENS:
            out = torch.empty(Bv, n, 64 * Cc, dtype=torch.float32, device=x.device)
        else:
            r = int(round(n ** 0.5))
            if r * r != n:
                raise ValueError(f"num_landm={n} is not a square number")
            out = torch.empty(Bv, Cc, 8 * r, 8 * r, dtype=torch.float32, device=x.device)
        _lib.call("lafs_gather_fwd", x.data_ptr(), th.data_ptr(), out.data_ptr(), Bv, Cc, H, W, n,
                  layout, coord_mode, _lib.stream())
        ctx.save_for_backward(x, th)
        ctx.cfg = (layout, coord_mode)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, th = ctx.saved_tensors
        layout, coord_mode = ctx.cfg
        Bv, Cc, H, W = x.shape
        n = th.shape[1]
        g = grad_out.detach().float().contiguous()
        need_img, need_th = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gi = torch.zeros_like(x) if need_img else None
        gt = torch.empty_like(th) if need_th else None
        _lib.call("lafs_gather_bwd", x.data_ptr(), th.data_ptr(), g.data_ptr(), _lib.ptr(gi), _lib.ptr(gt),
                  Bv, Cc, H, W, n, layout, coord_mode, _lib.stream())
        return gi, gt, None, None


def extract_patches_pytorch_gridsample(imgs, landmarks, patch_shape, num_landm=49):
    """Drop-in for the reference function.  `patch_shape` must describe 8x8 patches (the only
    value the reference ever passes: ViT_face.py:606,1264)."""
    ps = [int(v) for v in (patch_shape.tolist() if torch.is_tensor(patch_shape) else patch_shape)]
    if ps != [8, 8]:
        raise ValueError(f"only 8x8 patches are implemented (reference default), got {ps}")
    return _GatherFn.apply(imgs, landmarks[:, :num_landm], _lib.LAYOUT_MOSAIC, COORD_MODE)


def extract_tokens(imgs, landmarks, num_landm=None):
    """[B, n, 192] tokens in '(p1 p2 c)' feature order, without materialising the mosaic."""
    if num_landm is not None:
        landmarks = landmarks[:, :num_landm]
    return _GatherFn.apply(imgs, landmarks, _lib.LAYOUT_TOKENS, COORD_MODE)


class _LandmarkPostFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, raw, scale):
        x = raw.detach().float().contiguous()
        B, m = x.shape
        out = torch.empty(B, m // 2, 2, dtype=torch.float32, device=x.device)
        _lib.call("lafs_landmark_post", x.data_ptr(), None, None, out.data_ptr(), None, B, m // 2, 0,
                  float(scale), _lib.stream())
        ctx.save_for_backward(x)
        ctx.scale = float(scale)
        return out

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        B, m = x.shape
        g = g.detach().float().contiguous()
        gr = torch.empty_like(x)
        _lib.call("lafs_landmark_post_bwd", x.data_ptr(), g.data_ptr(), gr.data_ptr(), B, m // 2, ctx.scale,
                  _lib.stream())
        return gr, None


def landmark_post(raw, noise=None, extract_id=None, scale=111.0):
    """theta = (raw-min)/(max-min)*scale jointly per sample, view [B,n,2], + noise, gather.
    With noise / extract_id (the SSL, no-grad use) the fused kernel applies them; the plain
    form is differentiable w.r.t. raw (finetune path, ViT_face.py:694-706)."""
    _lib.require_cuda(raw, noise, extract_id)
    if raw.dim() != 2 or raw.shape[1] % 2:
        raise ValueError(f"raw must be [B, 2n], got {tuple(raw.shape)}")
    if noise is None and extract_id is None:
        return _LandmarkPostFn.apply(raw, scale)
    x = raw.detach().float().contiguous()
    B, m = x.shape
    n = m // 2
    nz = None if noise is None else noise.detach().float().contiguous()
    idx = None
    keep = 0
    if extract_id is not None:
        idx = extract_id.detach().reshape(B, -1).to(torch.int64).contiguous()
        keep = idx.shape[1]
    out = torch.empty(B, keep if idx is not None else n, 2, dtype=torch.float32, device=x.device)
    _lib.call("lafs_landmark_post", x.data_ptr(), _lib.ptr(nz), _lib.ptr(idx), out.data_ptr(), None,
              B, n, keep, float(scale), _lib.stream())
    return out


class PatchEmbedWeights:
    """bf16, K-permuted copy of one or two `patch_to_embedding = nn.Linear(192, dim)` layers
    (ViT_face.py:619) for the fused gather->embed kernel.  Two layers (student, teacher) share
    one gather of the patches.  Call refresh() after the weights change (optimizer / EMA step)."""

    def __init__(self, linears):
        self.linears = [(w, b) for (w, b) in linears]
        if not 1 <= len(self.linears) <= 2:
            raise ValueError("one or two (weight, bias) pairs expected")
        w0 = self.linears[0][0]
        _lib.require_cuda(w0)
        self.dim = w0.shape[0]
        if any(w.shape != (self.dim, 192) for w, _ in self.linears):
            raise ValueError("patch_to_embedding weights must be [dim, 192]")
        m = len(self.linears)
        self.w_perm = torch.empty(m * self.dim, 192, dtype=torch.bfloat16, device=w0.device)
        self.bias = torch.empty(m * self.dim, dtype=torch.float32, device=w0.device)
        self.refresh()

    @torch.no_grad()
    def refresh(self):
        for i, (w, b) in enumerate(self.linears):
            wd = w.detach().float().contiguous()
            bd = None if b is None else b.detach().float().contiguous()
            _lib.call("lafs_embed_weight_prep", wd.data_ptr(), _lib.ptr(bd), self.dim,
                      self.w_perm.data_ptr() + i * self.dim * 192 * 2,
                      self.bias.data_ptr() + i * self.dim * 4, _lib.stream())


def new_token_buffer(rows, device):
    """bf16 [rows, 208] buffer for the tokens `gather_embed(..., save_tokens=buf)` keeps for the weight gradient:
    columns 0..191 are written by the kernel on every call, column 192 is the ones column (bias gradient) and
    columns 193..207 stay zero -- set here once, so the buffer can be reused step after step."""
    buf = torch.empty(rows, _lib.TOK_LD, dtype=torch.bfloat16, device=device)
    buf[:, 192:] = 0
    buf[:, 192] = 1
    return buf


@torch.no_grad()
def embed_backward_weight(grad_emb, tokens_perm, grad_w=None, grad_b=None, accumulate=False):
    """grad of patch_to_embedding's weight [dim,192] and bias [dim] (fp32) from grad_emb [.., dim] (bf16) and the
    tokens the forward kept (new_token_buffer + gather_embed(save_tokens=)): one tcgen05 split-K GEMM."""
    _lib.require_cuda(grad_emb, tokens_perm)
    dim = grad_emb.shape[-1]
    g = grad_emb.detach().to(torch.bfloat16).contiguous().view(-1, dim)
    M = g.shape[0]
    if tokens_perm.shape != (M, _lib.TOK_LD) or tokens_perm.dtype != torch.bfloat16 or not tokens_perm.is_contiguous():
        raise ValueError(f"tokens_perm must be a contiguous bf16 [{M}, {_lib.TOK_LD}] buffer")
    dev = g.device
    if grad_w is None:
        grad_w, accumulate = torch.empty(dim, 192, dtype=torch.float32, device=dev), False
    if grad_b is None:
        grad_b = torch.empty(dim, dtype=torch.float32, device=dev) if not accumulate else None
    nbytes = _lib.lib().lafs_embed_bwd_workspace_bytes(M, dim)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.call("lafs_embed_bwd_weight_perm", g.data_ptr(), tokens_perm.data_ptr(), M, dim, grad_w.data_ptr(), _lib.ptr(grad_b),
              1 if accumulate else 0, ws.data_ptr(), nbytes, _lib.stream())
    return grad_w, grad_b


@torch.no_grad()
def gather_embed(imgs, landmarks, weights: PatchEmbedWeights, out_dtype=torch.bfloat16, mean=0.5, std=0.5, save_tokens=None,
                 seq=None, drop_p=0.0, seed=0):
    """tokens(imgs, landmarks) @ W^T + b for every model in `weights`, without materialising the
    patches: list of [B, n, dim] tensors.  Forward only (the SSL landmark CNN is frozen and
    the teacher has no gradient).

    imgs: fp32 [B,3,112,112] already normalised (the reference's tensors), or uint8 decoded pixels;
    for uint8 the reference's ToTensor + Normalize(mean, std) (lafs_train.py:800-803) runs inside
    the kernel, which quarters the image bytes moved over PCIe / read from HBM.
    save_tokens: a new_token_buffer(B*n) the kernel also fills with the gathered bf16 tokens (training path).
    seq: one (pos_embedding, cls_token) pair per model -> the kernel's epilogue also does ViT_face.py:762-768
    (`cat(cls, x)`, `+= pos_embedding[:, :n+1]`, dropout(drop_p)) and the outputs are the transformer inputs
    [B, n+1, dim]; `seed` selects the dropout mask (counter-based, not torch's stream)."""
    _lib.require_cuda(imgs, landmarks)
    if imgs.dtype == torch.uint8:
        x = imgs.detach().contiguous()
        in_dtype, scale, shift = _lib.U8, 1.0 / (255.0 * std), -mean / std
    else:
        x = imgs.detach().float().contiguous()
        in_dtype, scale, shift = _lib.F32, 1.0, 0.0
    th = landmarks.detach().float().contiguous()
    Bv, Cc, H, W = x.shape
    if Cc != 3:
        raise ValueError("the fused path expects 3-channel images")
    n = th.shape[1]
    m = len(weights.linears)
    pc = [None, None, None, None]
    if seq is not None:
        if len(seq) != m:
            raise ValueError("seq needs one (pos_embedding, cls_token) pair per model")
        for i, (pos, cls) in enumerate(seq):
            _lib.require_cuda(pos, cls)
            pos2 = pos.detach().float().reshape(-1, weights.dim).contiguous()
            cls1 = cls.detach().float().reshape(-1).contiguous()
            if pos2.shape[0] < n + 1 or cls1.numel() != weights.dim:
                raise ValueError(f"pos_embedding needs >= {n + 1} rows of {weights.dim}, cls_token {weights.dim} values")
            pc[2 * i], pc[2 * i + 1] = pos2, cls1
    rows = n + 1 if seq is not None else n
    outs = [torch.empty(Bv, rows, weights.dim, dtype=out_dtype, device=x.device) for _ in range(m)]
    if save_tokens is not None and (save_tokens.shape != (Bv * n, _lib.TOK_LD) or save_tokens.dtype != torch.bfloat16
                                    or not save_tokens.is_contiguous()):
        raise ValueError(f"save_tokens must be a contiguous bf16 [{Bv * n}, {_lib.TOK_LD}] buffer (new_token_buffer)")
    _lib.call("lafs_gather_embed_seq_fwd", x.data_ptr(), in_dtype, scale, shift, th.data_ptr(),
              weights.w_perm.data_ptr(), weights.bias.data_ptr(), outs[0].data_ptr(),
              outs[1].data_ptr() if m > 1 else None, _lib.dtype_code(outs[0]), Bv, H, W, n, weights.dim, m,
              _lib.ptr(save_tokens), _lib.ptr(pc[0]), _lib.ptr(pc[1]), _lib.ptr(pc[2]), _lib.ptr(pc[3]),
              float(drop_p), int(seed) & 0xFFFFFFFF, _lib.stream())
    return outs


class _GatherEmbedTrainFn(torch.autograd.Function):
    """Differentiable gather -> patch_to_embedding for the finetune path (the landmark CNN and the
    embedding are trained, ViT_face.py:706,760-761).  Forward: the fused tcgen05 kernel, which also keeps the
    gathered tokens in bf16 (14 % of the output bytes; no fp32 token tensor, no re-gather).  Backward:
        [grad_weight | grad_bias] = grad_emb^T . [tokens | 1]      (one split-K GEMM, lafs_embed_bwd_weight_perm)
        grad_tokens = grad_emb . W  ->  lafs_gather_bwd  ->  grad_landmarks (, grad_imgs)
    on the same tcgen05 GEMM the margin head's backward uses (bf16 operands, fp32 accumulation)."""

    @staticmethod
    def forward(ctx, imgs, theta, weight, bias):
        w = PatchEmbedWeights([(weight, bias)])
        need_w = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        tok = new_token_buffer(theta.shape[0] * theta.shape[1], theta.device) if need_w else None
        (emb,) = gather_embed(imgs, theta, w, out_dtype=torch.bfloat16, save_tokens=tok)
        ctx.save_for_backward(imgs, theta, weight)
        ctx.tok = tok
        ctx.has_bias = bias is not None
        return emb

    @staticmethod
    def backward(ctx, grad_emb):
        imgs, theta, weight = ctx.saved_tensors
        need_img, need_th, need_w, need_b = ctx.needs_input_grad
        x = imgs.detach().float().contiguous()
        th = theta.detach().float().contiguous()
        Bv, Cc, H, W = x.shape
        n = th.shape[1]
        dim = weight.shape[0]
        M = Bv * n
        g = grad_emb.detach().to(torch.bfloat16).contiguous().view(M, dim)
        dev = g.device
        gw = gb = gi = gt = None
        if (need_w or need_b) and ctx.tok is not None:
            gw, gb = embed_backward_weight(g, ctx.tok)
            gw = gw.to(weight.dtype) if need_w else None
            gb = gb if (need_b and ctx.has_bias) else None
        if need_img or need_th:
            w16 = weight.detach().to(torch.bfloat16).contiguous()
            gtok = torch.empty(M, 64 * Cc, dtype=torch.float32, device=dev)
            _lib.call("lafs_embed_bwd_tokens", g.data_ptr(), w16.data_ptr(), M, dim, gtok.data_ptr(), _lib.stream())
            gi = torch.zeros_like(x) if need_img else None
            gt = torch.empty_like(th) if need_th else None
            _lib.call("lafs_gather_bwd", x.data_ptr(), th.data_ptr(), gtok.data_ptr(), _lib.ptr(gi), _lib.ptr(gt),
                      Bv, Cc, H, W, n, _lib.LAYOUT_TOKENS, COORD_MODE, _lib.stream())
        return gi, gt, gw, gb


def gather_embed_train(imgs, landmarks, weight, bias=None):
    """tokens(imgs, landmarks) @ weight^T + bias -> [B, n, dim] bf16, differentiable w.r.t. landmarks,
    imgs, weight and bias (fp32 [B,3,112,112] images; weight [dim,192], dim % 128 == 0, n <= 208)."""
    _lib.require_cuda(imgs, landmarks, weight, bias)
    if imgs.dim() != 4 or imgs.shape[1] != 3 or imgs.dtype == torch.uint8:
        raise ValueError("gather_embed_train expects fp32 [B,3,H,W] images")
    return _GatherEmbedTrainFn.apply(imgs, landmarks, weight, bias)
