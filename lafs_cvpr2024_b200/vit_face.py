"""Part-fViT and landmark-CNN wrappers with the reference's constructor / forward surface and
state-dict keys, routed through the sm_100a kernels for the hot-path pieces.

    ViT_face_landmark_patch8(...)        face_pre_pro/ViT_face.py:560-795
    face_landmark_4simmin_glo_loc(...)   face_pre_pro/ViT_face.py:1218-1409

What runs where
  * landmark tail (joint min-max, noise, re-sampling)            -> lafs_landmark_post   (kernel)
  * patch extraction / token layout                              -> lafs_gather_fwd/_bwd (kernel)
  * extraction + patch_to_embedding without gradients            -> lafs_gather_embed_fwd (tcgen05)
  * margin head `self.loss`                                      -> CosFace / ArcFace    (tcgen05)
  * transformer blocks, LayerNorm, MobileNetV3 landmark trunk    -> stock PyTorch, as in the
    reference (outside the hot path); the trunk (landmark_trunk.py) keeps the
    reference's module tree / checkpoint keys, and another module can be injected via `stn=`.

Checkpoint compatibility: parameter / buffer names equal the reference's (`pos_embedding`,
`patch_to_embedding.{weight,bias}`, `cls_token`, `transformer.layers.{i}.{0,1}.fn.…`,
`mlp_head.0.*`, `loss.weight`, `stn.*`, `output_layer.1.*`, `global_token.1.*`, `mask_token`).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .margin_head import ArcFace, CosFace
from .patches import PatchEmbedWeights, extract_tokens, gather_embed, gather_embed_train, landmark_post

MIN_NUM_PATCHES = 15  # ViT_face.py:21


def _default_trunk():
    """The reference's `MobileNetV3_backbone(mode='large')` (ViT_face.py:611,1269): same module tree and
    checkpoint keys, restated in landmark_trunk.py (stock PyTorch convolutions, outside the hot path)."""
    from .landmark_trunk import MobileNetV3LargeTrunk
    return MobileNetV3LargeTrunk()


class DropPath(nn.Module):
    """Per-sample stochastic depth."""

    def __init__(self, p=0.0):
        super().__init__()
        self.drop_prob = p

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.dim() - 1)).bernoulli_(keep)
        return x * mask / keep


class Residual_droppath(nn.Module):
    def __init__(self, fn, drop_path_rate=0.1):
        super().__init__()
        self.fn = fn
        self.drop_path = DropPath(drop_path_rate) if drop_path_rate > 0.0 else nn.Identity()

    def forward(self, x, **kw):
        return self.drop_path(self.fn(x, **kw)) + x


class PreNorm(nn.Module):
    def __init__(self, dim, fn):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.fn = fn

    def forward(self, x, **kw):
        return self.fn(self.norm(x), **kw)


class FeedForward(nn.Module):
    def __init__(self, dim, hidden_dim, dropout=0.0):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, hidden_dim), nn.GELU(), nn.Dropout(dropout),
                                 nn.Linear(hidden_dim, dim), nn.Dropout(dropout))

    def forward(self, x):
        return self.net(x)


class Attention(nn.Module):
    """Multi-head attention with the reference's quirks: heads*dim_head need not equal
    dim and the logits are scaled by dim**-0.5, not dim_head**-0.5."""

    def __init__(self, dim, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner = dim_head * heads
        self.heads = heads
        self.scale = dim ** -0.5
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, dim), nn.Dropout(dropout))
        self.attention_score = 0

    # The reference stores softmax(q k^T * scale) of EVERY forward in `attention_score` (ViT_face.py:178); its only
    # reader is the evaluation-time attention-map plot (util/utils.py:662).  The flash path never forms that
    # [b, h, n, n] tensor, so the last forward's q and k are kept (two views of the qkv projection, no copy) and
    # the scores are computed when somebody reads the attribute.
    @property
    def attention_score(self):
        qk = self.__dict__.get("_qk")
        if qk is not None:
            q, k = qk
            return (torch.einsum('bhid,bhjd->bhij', q, k) * self.scale).softmax(dim=-1)
        return self.__dict__.get("_attention_score", 0)

    @attention_score.setter
    def attention_score(self, value):
        self.__dict__["_qk"] = None
        self.__dict__["_attention_score"] = value

    def forward(self, x, mask=None):
        b, n, _ = x.shape
        q, k, v = (t.view(b, n, self.heads, -1).transpose(1, 2) for t in self.to_qkv(x).chunk(3, dim=-1))
        if mask is None:
            out = F.scaled_dot_product_attention(q, k, v, scale=self.scale)
            self.__dict__["_qk"] = (q.detach(), k.detach())
        else:
            # the reference's masked form, op for op (ViT_face.py:164-176): masked logits are REPLACED by
            # -finfo.max, so a fully masked query row attends uniformly and stays finite (a boolean SDPA mask
            # would give NaN / 0 there); `attention_score` is kept like the reference does
            dots = torch.einsum('bhid,bhjd->bhij', q, k) * self.scale
            m = F.pad(mask.flatten(1), (1, 0), value=True)
            assert m.shape[-1] == dots.shape[-1], 'mask has incorrect dimensions'
            m = m[:, None, :] * m[:, :, None]
            dots = dots.masked_fill(~m[:, None], -torch.finfo(dots.dtype).max)
            attn = dots.softmax(dim=-1)
            self.attention_score = attn.detach()
            out = torch.einsum('bhij,bhjd->bhid', attn, v)
        return self.to_out(out.transpose(1, 2).reshape(b, n, -1))


class Transformer(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, mlp_dim, dropout):
        super().__init__()
        self.layers = nn.ModuleList([
            nn.ModuleList([Residual_droppath(PreNorm(dim, Attention(dim, heads, dim_head, dropout))),
                           Residual_droppath(PreNorm(dim, FeedForward(dim, mlp_dim, dropout)))])
            for _ in range(depth)])

    def forward(self, x, mask=None):
        for attn, ff in self.layers:
            x = ff(attn(x, mask=mask))
        return x


def _tokens_and_embedding(imgs, theta, linear, training_path, allow_tc=True):
    """patch tokens -> patch_to_embedding on the fused tcgen05 gather->embed kernel (no token tensor in
    HBM).  With gradients its backward runs on the tcgen05 GEMMs of patches.gather_embed_train; shapes
    the fused kernel does not cover (dim % 128 != 0, more than 208 landmarks) take the differentiable
    fp32 gather kernel + nn.Linear -- and so does a model that asked for full precision (allow_tc=False: fp16=False
    in the constructor and no autocast region): the tensor-core path rounds tokens and weights to bf16."""
    fused_ok = (allow_tc and linear.weight.shape[0] % 128 == 0 and theta.shape[1] <= 208 and imgs.shape[1] == 3
                and tuple(imgs.shape[-2:]) == (112, 112))
    if training_path:
        if fused_ok:
            return gather_embed_train(imgs, theta, linear.weight, linear.bias).to(linear.weight.dtype)
        tok = extract_tokens(imgs, theta)
        return linear(tok.to(linear.weight.dtype))
    if not fused_ok:
        return linear(extract_tokens(imgs, theta).to(linear.weight.dtype))
    w = PatchEmbedWeights([(linear.weight, linear.bias)])
    (emb,) = gather_embed(imgs, theta, w, out_dtype=torch.bfloat16)
    return emb.to(linear.weight.dtype)


def image_grid_tokens(x, p):
    """einops 'b c (h p1) (w p2) -> b (h w) (p1 p2 c)' (ViT_face.py:759) for image input without landmarks."""
    b, c, H, W = x.shape
    return x.reshape(b, c, H // p, p, W // p, p).permute(0, 2, 4, 3, 5, 1).reshape(b, (H // p) * (W // p), p * p * c)


def standard_grid_tokens(x, num_land, random_prob, shuffle, extract_mosaic):
    """`use_standcoord=True` of the reference (ViT_face.py:717-746): patches at the fixed grid centres
    4, 12, ..., (+ randn*3 with Random_prob, re-sampled with replacement with shuffle; both drawn on the CPU
    default generator in the reference's order), the mosaic transposed (`x.permute(0,1,3,2)`, :746) and
    re-laid-out as '(h w) (p1 p2 c)' tokens (:759).  `extract_mosaic(imgs, theta)` is the patch extractor
    (the sm_100a gather kernel in the product; the tests pass the oracle's to check this composition on CPU
    against the reference).  Returns (theta [b,n,2], tokens [b,n,192])."""
    b = x.shape[0]
    rc = torch.arange(0, math.sqrt(num_land)) * 8 + 4              # float32, like the reference's torch.arange(0, np.sqrt(n))
    cx, cy = torch.meshgrid(rc, rc, indexing="ij")
    theta = torch.stack((cx, cy), 2).view(1, -1, 2).repeat(b, 1, 1).to(x.device)
    n = theta.shape[1]
    if random_prob:
        theta = theta + (torch.randn(theta.shape) * 3).to(x.device)
    if shuffle:
        idx = torch.randint(0, n, (b, n, 1)).to(x.device).repeat(1, 1, 2)
        theta = torch.gather(theta, 1, idx)
    theta = theta[:, :num_land]
    m = extract_mosaic(x, theta).permute(0, 1, 3, 2)                # [b, c, 8r, 8r], height and width swapped
    c, r = m.shape[1], m.shape[2] // 8
    tokens = m.reshape(b, c, r, 8, r, 8).permute(0, 2, 4, 3, 5, 1).reshape(b, r * r, 64 * c)
    return theta, tokens


class ViT_face_landmark_patch8(nn.Module):
    def __init__(self, *, loss_type, GPU_ID, num_class, image_size, patch_size, dim, depth, heads, mlp_dim,
                 pool='cls', num_patches=None, channels=3, dim_head=64, dropout=0., emb_dropout=0., fp16=True,
                 with_land=False, use_standcoord=False, Random_prob=False, shuffle=False, stn=None):
        super().__init__()
        if num_patches is None:
            num_patches = (image_size // patch_size) ** 2
        patch_dim = channels * patch_size ** 2
        assert num_patches > MIN_NUM_PATCHES, f'your number of patches ({num_patches}) is way too small for attention to be effective (at least 16). Try decreasing your patch size'
        assert pool in {'cls', 'mean'}, 'pool type must be either cls (cls token) or mean (mean pooling)'
        if patch_size != 8:
            raise ValueError("the sm_100a gather kernels implement the reference's 8x8 patches")
        if use_standcoord and with_land:
            raise NotImplementedError("with_land and use_standcoord together (the reference would re-extract patches "
                                      "from the landmark mosaic) is not built")
        self.use_standcoord = use_standcoord
        self.patch_size = patch_size
        self.fp16 = fp16
        self.num_patches = num_patches
        self.row_num = int(math.sqrt(num_patches))
        self.with_land = with_land
        if with_land:
            self.stn = stn if stn is not None else _default_trunk()
            self.output_layer = nn.Sequential(nn.Dropout(p=0.5), nn.Linear(160, self.row_num * self.row_num * 2))
        self.patch_shape = torch.tensor([patch_size, patch_size])
        self.theta = 0
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + 1, dim))
        self.patch_to_embedding = nn.Linear(patch_dim, dim)
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.dropout = nn.Dropout(emb_dropout)
        self.transformer = Transformer(dim, depth, heads, dim_head, mlp_dim, dropout)
        self.pool = pool
        self.to_latent = nn.Identity()
        self.mlp_head = nn.Sequential(nn.LayerNorm(dim))
        self.loss_type = loss_type
        self.GPU_ID = GPU_ID
        self.Random_prob = Random_prob
        self.shuffle = shuffle
        if loss_type == 'CosFace':
            self.loss = CosFace(in_features=dim, out_features=num_class, device_id=GPU_ID, m=0.4)
        elif loss_type == 'ArcFace':
            self.loss = ArcFace(in_features=dim, out_features=num_class, device_id=GPU_ID)
        elif loss_type != 'None':
            raise ValueError(f"loss_type {loss_type!r}: the reference defines only CosFace (ArcFace is provided "
                             "here; Softmax / SFace are undefined in the reference as well)")

    def landmarks(self, x):
        """[B,n,2] landmark coordinates of a face batch (ViT_face.py:680-706), differentiable."""
        feat = self.stn(x).mean(dim=(-2, -1))
        theta = landmark_post(self.output_layer(feat))
        self.theta = theta
        return theta

    def _num_land(self, x):
        """ViT_face.py:663-677: (H/p)^2 landmarks, except the 144-patch model on 112-pixel faces."""
        if self.num_patches == 144 and x.shape[-2] == 112:
            return self.num_patches
        return (x.shape[-2] // self.patch_size) ** 2

    def forward(self, x, label=None, mask=None, visualize=False, save_token=False, opt=None, keep_num=None,
                glo_diff=False):
        theta = None
        if x.dim() == 4 and not self.with_land:
            # images without the landmark CNN (ViT_face.py:717-761): fixed-grid patches (use_standcoord) or the
            # plain ViT patch grid; stock PyTorch re-layout around the gather kernel
            num_land = self._num_land(x)
            if self.use_standcoord:
                _lib.require_cuda(x)
                from .patches import extract_patches_pytorch_gridsample
                theta, tok = standard_grid_tokens(
                    x.float(), num_land, self.Random_prob, self.shuffle,
                    lambda im, th: extract_patches_pytorch_gridsample(im, th, self.patch_shape, th.shape[1]))
            else:
                tok = image_grid_tokens(x, self.patch_size)
            x = self.patch_to_embedding(tok.to(self.patch_to_embedding.weight.dtype))
        elif x.dim() == 4:
            _lib.require_cuda(x)
            theta = self.landmarks(x.float())
            num_land = self._num_land(x)
            need_grad = torch.is_grad_enabled() and (theta.requires_grad or x.requires_grad
                                                     or self.patch_to_embedding.weight.requires_grad)
            lin = self.patch_to_embedding
            # half-precision contract of the reference: its models are built with fp16=True and trained under autocast
            # (ViT_face.py:561, train_largescale.py:803-804); only then does patch_to_embedding run on bf16 operands
            use_tc = bool(self.fp16) or torch.is_autocast_enabled()
            if (use_tc and not need_grad and lin.weight.shape[0] % 128 == 0 and num_land <= 208 and x.shape[1] == 3
                    and tuple(x.shape[-2:]) == (112, 112) and self.pos_embedding.shape[1] >= num_land + 1):
                # inference / frozen path: cls row, pos_embedding and dropout are part of the fused kernel's epilogue
                # (SURVEY 8f row 2): the transformer input leaves ONE kernel
                p_drop = self.dropout.p if self.training else 0.0
                seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item()) if p_drop > 0 else 0
                (x,) = gather_embed(x.float(), theta[:, :num_land], PatchEmbedWeights([(lin.weight, lin.bias)]),
                                    seq=[(self.pos_embedding, self.cls_token)], drop_p=p_drop, seed=seed)
                return self._after_embedding(x.to(lin.weight.dtype), label, mask, visualize, save_token, opt, theta)
            x = _tokens_and_embedding(x.float(), theta[:, :num_land], self.patch_to_embedding, need_grad, use_tc)
        else:
            x = self.patch_to_embedding(x)                       # SSL path: tokens prepared by the landmark CNN
        b, n, _ = x.shape
        x = torch.cat((self.cls_token.expand(b, -1, -1).to(x.dtype), x), dim=1)
        x = x + self.pos_embedding[:, :(n + 1)].to(x.dtype)
        x = self.dropout(x)
        return self._after_embedding(x, label, mask, visualize, save_token, opt, theta)

    def _after_embedding(self, x, label, mask, visualize, save_token, opt, theta):
        x = self.transformer(x, mask)
        tokens = x[:, 1:] if save_token else None
        x = x.mean(dim=1) if self.pool == 'mean' else x[:, 0]
        emb = self.mlp_head(self.to_latent(x))
        if save_token:
            return emb, tokens, self.theta
        if label is not None:
            return self.loss(emb.float(), label), (emb if opt is not None else self.theta)
        return (emb, theta) if visualize else emb

    def forward_loss(self, x, label, label_b=None, lam=1.0):
        """Fused finetune step tail: embedding -> margin head -> cross-entropy, no [B,C] logits."""
        emb = self.forward(x)
        return self.loss.forward_loss(emb.float(), label, label_b, lam)


class face_landmark_4simmin_glo_loc(nn.Module):
    def __init__(self, *, loss_type, GPU_ID, num_class, image_size, patch_size, dim, depth, heads, mlp_dim,
                 pool='cls', num_patches=None, channels=3, dim_head=64, dropout=0., emb_dropout=0., fp16=True,
                 stn=None):
        super().__init__()
        if num_patches is None:
            num_patches = (image_size // patch_size) ** 2
        patch_dim = channels * patch_size ** 2
        assert num_patches > MIN_NUM_PATCHES, f'your number of patches ({num_patches}) is way too small for attention to be effective (at least 16). Try decreasing your patch size'
        if patch_size != 8:
            raise ValueError("the sm_100a gather kernels implement the reference's 8x8 patches")
        self.patch_size = patch_size
        self.fp16 = fp16
        self.num_patches = num_patches
        self.row_num = int(math.sqrt(num_patches))
        self.stn = stn if stn is not None else _default_trunk()
        self.dim = dim
        self.output_layer = nn.Sequential(nn.Dropout(p=0.5), nn.Linear(160, self.row_num * self.row_num * 2))
        self.global_token = nn.Sequential(nn.Dropout(p=0.5), nn.Linear(160, dim))
        self.patch_shape = torch.tensor([patch_size, patch_size])
        self.theta = 0
        self.pos_embedding = nn.Parameter(torch.randn(1, num_patches + 1, dim))
        self.patch_to_embedding = nn.Linear(patch_dim, dim)
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.dropout = nn.Dropout(emb_dropout)
        self.loss_type = loss_type
        self.GPU_ID = GPU_ID
        self.num_features = dim
        self.in_chans = channels
        self.mask_token = nn.Parameter(torch.zeros(1, 1, dim))
        nn.init.trunc_normal_(self.mask_token, std=.02, a=-.02, b=.02)

    def landmarks(self, x, Random_prob=False, return_prob=False, ran_sample=False):
        """theta as the reference computes it (ViT_face.py:1332-1380).  The noise and the re-sampled
        indices are drawn on the CPU default generator in the reference's order."""
        feat = self.stn(x).mean(dim=(-2, -1))
        raw = self.output_layer(feat)
        n = self.row_num * self.row_num
        noise = idx = None
        if Random_prob:
            noise = (torch.randn(raw.shape[0], n, 2) * 5).to(raw.device)
            if not return_prob:
                keep = 36 if ran_sample else n
                idx = torch.randint(0, n, (raw.shape[0], keep, 1)).to(raw.device)
        return landmark_post(raw, noise, idx)

    def forward(self, x, x_Aug=None, keep_num=None, patch_shape=None, Random_prob=False, return_prob=False,
                ran_sample=False, random_coor=False, return_land=False):
        _lib.require_cuda(x)
        if random_coor:
            num_land = 25 if ran_sample else self.row_num * self.row_num
            theta = (torch.rand(x.shape[0], num_land, 2) * 111.0).to(x.device)
        else:
            theta = self.landmarks(x, Random_prob, return_prob, ran_sample)
            self.theta = theta
            num_land = self._keep_num(x.shape[-2], theta.shape[1], Random_prob and not return_prob)
        if return_land:
            return theta, x
        src = x if x_Aug is None else x_Aug
        from .patches import extract_patches_pytorch_gridsample
        return theta, extract_patches_pytorch_gridsample(src, theta[:, :num_land], self.patch_shape, num_land)

    def _keep_num(self, imgshape, n_theta, resampled):
        """How many of theta's landmarks the reference hands to the extractor (ViT_face.py:1319-1325,1363-1385):
        all of a re-sampled set (36 or row_num^2); otherwise num_patches for the 144 / 196-patch models on 112-pixel
        faces and (H / p)^2 -- the FIRST that many -- for any other input size."""
        if resampled:
            return n_theta
        if self.num_patches in (144, 196) and imgshape == 112:
            return self.num_patches
        return (imgshape // self.patch_size) ** 2

    def forward_tokens(self, x, x_Aug=None, Random_prob=False, return_prob=False, ran_sample=False):
        """theta and the '(p1 p2 c)' token tensor [B,n,192] directly (fuses lafs_train.py:535-538)."""
        theta = self.landmarks(x, Random_prob, return_prob, ran_sample)
        self.theta = theta
        return theta, extract_tokens(x if x_Aug is None else x_Aug, theta)
