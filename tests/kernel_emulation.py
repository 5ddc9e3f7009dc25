"""numpy restatement of the *CUDA kernel's* fp32 op sequence for the bilinear landmark
gather (lafs_cvpr2024_b200/csrc/gather.cu), used on CPU to show that the sequence is
bit-identical to the reference before any GPU time is spent.

Sequence per sample point (x shown; y identical), H = image side:
    p  = theta + (i - 4)                 fadd.rn
    g  = p / (H/2) - 1                   div.rn (or mul by fp32(2/H) in recip mode), fsub.rn
    a  = g + 1                           fadd.rn
    ix = fma(a, H/2, -0.5)               == ((g+1)*H - 1)/2 of grid_sampler_unnormalize
    x0 = floor(ix); w = ix - x0; e = 1 - w
    nw = s*e, ne = s*w, sw = n*e, se = n*w
    out = fma(v_se, se, fma(v_sw, sw, fma(v_ne, ne, v_nw*nw)))
"""
import numpy as np

f32 = np.float32


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def gather_tokens(imgs, th, recip=False):
    """imgs [B,C,H,W] fp32, th [B,n,2] fp32 -> [B,n,8,8,C] (i, j, c) fp32."""
    B, C, H, W = imgs.shape
    n = th.shape[1]
    half = f32(H * 0.5)
    off = np.arange(-4, 4, dtype=f32)
    out = np.zeros((B, n, 8, 8, C), f32)

    def coord(p):
        if recip:
            g = (p * f32(1.0 / float(half))).astype(f32) - f32(1)
        else:
            g = (p / half).astype(f32) - f32(1)
        a = (g.astype(f32) + f32(1)).astype(f32)
        return _fma(a, np.full_like(a, half), np.full_like(a, -0.5))

    ix = coord((th[:, :, 0:1] + off[None, None, :]).astype(f32))
    iy = coord((th[:, :, 1:2] + off[None, None, :]).astype(f32))
    x0, y0 = np.floor(ix), np.floor(iy)
    w = (ix - x0).astype(f32); e = (f32(1) - w).astype(f32)
    nn = (iy - y0).astype(f32); s = (f32(1) - nn).astype(f32)
    x0, y0 = x0.astype(np.int64), y0.astype(np.int64)

    def fetch(b, xx, yy):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        v = imgs[b][:, np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)]
        return np.where(ok[None], v, f32(0))

    for b in range(B):
        X0 = np.broadcast_to(x0[b][:, :, None], (n, 8, 8)); Y0 = np.broadcast_to(y0[b][:, None, :], (n, 8, 8))
        E = np.broadcast_to(e[b][:, :, None], (n, 8, 8)); Wt = np.broadcast_to(w[b][:, :, None], (n, 8, 8))
        S = np.broadcast_to(s[b][:, None, :], (n, 8, 8)); N = np.broadcast_to(nn[b][:, None, :], (n, 8, 8))
        nw = (S * E).astype(f32); ne = (S * Wt).astype(f32); sw = (N * E).astype(f32); se = (N * Wt).astype(f32)
        vnw = fetch(b, X0, Y0); vne = fetch(b, X0 + 1, Y0); vsw = fetch(b, X0, Y0 + 1); vse = fetch(b, X0 + 1, Y0 + 1)
        bc = lambda a: np.broadcast_to(a[None], vnw.shape)
        r = (vnw * nw[None]).astype(f32)
        r = _fma(vne, bc(ne), r); r = _fma(vsw, bc(sw), r); r = _fma(vse, bc(se), r)
        out[b] = np.transpose(r, (1, 2, 3, 0))
    return out


def xchg_slices(n4, world, nctas=64, threads=256, unroll=4):
    """Index sets (in float4 units) that rank r's CTAs cover in xchg_allreduce_kernel (csrc/exchange.cu):
    rank r owns [r*per, min(n4, (r+1)*per)), per = ceil(n4 / world); a thread starts at
    r_lo + (cta*threads + tid)*unroll and strides by nctas*threads*unroll."""
    per = -(-n4 // world)
    out = []
    for r in range(world):
        lo, hi = r * per, min(n4, (r + 1) * per)
        idx = []
        for cta in range(nctas):
            for tid in range(threads):
                base = lo + (cta * threads + tid) * unroll
                while base < hi:
                    idx.extend(i for i in range(base, base + unroll) if i < hi)
                    base += nctas * threads * unroll
        out.append(sorted(idx))
    return out


def xchg_allreduce(parts):
    """Two-shot all-reduce as the kernel performs it: rank r sums slice r of every rank's partial in rank
    order (fp32, fixed order), every rank receives every slice.  parts: list of equal-shape fp32 arrays
    whose size is a multiple of 4.  Returns the array every rank ends with."""
    world = len(parts)
    flat = [np.asarray(p, dtype=np.float32).reshape(-1, 4) for p in parts]
    n4 = flat[0].shape[0]
    out = np.empty_like(flat[0])
    per = -(-n4 // world)
    for r in range(world):
        lo, hi = r * per, min(n4, (r + 1) * per)
        acc = np.zeros((max(hi - lo, 0), 4), dtype=np.float32)
        for q in range(world):
            acc = (acc + flat[q][lo:hi]).astype(np.float32)
        out[lo:hi] = acc
    return out.reshape(np.asarray(parts[0]).shape)
