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


def _merge_records(rec_t, rec_s, group):
    """merge_cta_records (csrc/dino.cu): the records of `group` adjacent slices -> one record of the same format
    (teacher: max, Z, A rescaled to the group max; student: max, sum)."""
    B, nsl = rec_t.shape[:2]
    ng = -(-nsl // group)
    out_t = np.zeros((B, ng) + rec_t.shape[2:])
    out_s = np.zeros((B, ng) + rec_s.shape[2:])
    for g in range(ng):
        rt, rs = rec_t[:, g * group:(g + 1) * group], rec_s[:, g * group:(g + 1) * group]
        m = rt[..., 0].max(1)
        f = np.exp2(rt[..., 0] - m[:, None])
        out_t[:, g, :, 0] = m
        out_t[:, g, :, 1] = (rt[..., 1] * f).sum(1)
        out_t[:, g, :, 2] = (rt[..., 2] * f).sum(1)
        m = rs[..., 0].max(1)
        out_s[:, g, :, 0] = m
        out_s[:, g, :, 1] = (rs[..., 1] * np.exp2(rs[..., 0] - m[:, None])).sum(1)
    return out_t, out_s


def dino_sliced_loss(student, teacher, center, ncrops, inv_ts, inv_tt, slice_cols=256, cta_slices=1):
    """The DINO forward as the kernels decompose it (csrc/dino.cu), in float64: per (sample, column slice)
    partial records in the log2 domain -- teacher view iq: (max, Z, A) with A = sum_k e_iq,k * (S_k - s_iq,k),
    S = sum of the student rows; student crop v: (max, sum) -- then dino_finish's two-pass merge over the
    slices, the per-sample loss  sum_v n_v*lse(s_v/ts) - (1/ts) * sum_iq A_iq/Z_iq  and the mean over
    samples / (2*ncrops - 2).  cta_slices = 8 adds the per-CTA pre-merge of 8 adjacent slice records in between
    (merge_cta_records).  Returns (loss, per-row log2-domain lse [ncrops+2, B], column sums [K])."""
    s = np.asarray(student, dtype=np.float64)
    t = np.asarray(teacher, dtype=np.float64)
    c = np.asarray(center, dtype=np.float64).reshape(-1)
    K = s.shape[1]
    B = s.shape[0] // ncrops
    log2e = 1.4426950408889634
    a_s, a_t = inv_ts * log2e, inv_tt * log2e
    nsl = -(-K // slice_cols)
    rec_t = np.zeros((B, nsl, 2, 3))
    rec_s = np.zeros((B, nsl, ncrops, 2))
    colsum = t.sum(0)
    for b in range(B):
        for sl in range(nsl):
            k0, k1 = sl * slice_cols, min(K, (sl + 1) * slice_cols)
            srows = np.stack([s[v * B + b, k0:k1] for v in range(ncrops)])
            S = srows.sum(0)
            for iq in range(2):
                x = t[iq * B + b, k0:k1] * a_t - c[k0:k1] * a_t
                mx = x.max()
                e = np.exp2(x - mx)
                rec_t[b, sl, iq] = (mx, e.sum(), (e * (S - srows[iq])).sum())
            for v in range(ncrops):
                mx = srows[v].max() * a_s
                rec_s[b, sl, v] = (mx, np.exp2(srows[v] * a_s - mx).sum())
    if cta_slices > 1:      # the streaming kernel merges the records of a CTA's adjacent slices before they leave
        rec_t, rec_s = _merge_records(rec_t, rec_s, cta_slices)
    stats = np.zeros((ncrops + 2, B))
    total = 0.0
    for b in range(B):
        loss = 0.0
        for v in range(ncrops):
            m = rec_s[b, :, v, 0].max()
            z = (rec_s[b, :, v, 1] * np.exp2(rec_s[b, :, v, 0] - m)).sum()
            l2 = m + np.log2(z)
            stats[v, b] = l2
            loss += (1.0 if v < 2 else 2.0) * l2
        loss *= 0.6931471805599453
        for iq in range(2):
            m = rec_t[b, :, iq, 0].max()
            f = np.exp2(rec_t[b, :, iq, 0] - m)
            z, a = (rec_t[b, :, iq, 1] * f).sum(), (rec_t[b, :, iq, 2] * f).sum()
            stats[ncrops + iq, b] = m + np.log2(z)
            loss -= inv_ts * (a / z)
        total += loss
    return total / ((2 * ncrops - 2) * B), stats, colsum


def head_sharded_loss(cos, label, s, m, world, chunk=256):
    """CosFace + cross-entropy as the head kernels decompose it (csrc/head.cu, exchange.cu), in float64:
    per (row, rank, class chunk) online-softmax records (max in the log2 domain, sum-exp, target logit),
    merged per rank, then across ranks in rank order; classes are owned with torch.chunk's rule."""
    cos = np.asarray(cos, dtype=np.float64)
    B, C = cos.shape
    log2e = 1.4426950408889634
    step = -(-C // world)
    recs = []
    for r in range(world):
        lo, hi = min(r * step, C), min((r + 1) * step, C)
        mrun = np.full(B, -np.inf); lrun = np.zeros(B); tgt = np.zeros(B)
        for c0 in range(lo, hi, chunk):
            c1 = min(hi, c0 + chunk)
            z = cos[:, c0:c1].copy()
            for b in range(B):
                if c0 <= label[b] < c1:
                    z[b, label[b] - c0] -= m
                    tgt[b] = s * z[b, label[b] - c0]
            pm = z.max(1) * s * log2e
            mn = np.maximum(mrun, pm)
            lrun = lrun * np.exp2(mrun - mn) + np.exp2(z * s * log2e - mn[:, None]).sum(1)
            mrun = mn
        recs.append((mrun, lrun, tgt))
    M = np.full(B, -np.inf); L = np.zeros(B); T = np.zeros(B)
    for mr, lr, tg in recs:                      # rank order, as xchg_stats_kernel / head_merge_kernel
        ok = np.isfinite(mr)
        mn = np.where(ok, np.maximum(M, mr), M)
        L = np.where(ok, L * np.exp2(M - mn) + lr * np.exp2(np.where(ok, mr, 0.0) - mn), L)
        M = mn
        T += tg
    lse = (M + np.log2(L)) * 0.6931471805599453
    return float((lse - T).mean())


# ---- (f1) fused DINO head: the decomposition of csrc/dino_head.cu + dino_head.py, step by step ----------------
def _bf16(t):
    return t.bfloat16().float()


def dino_head_fused(xs, xt, vs, gs, vt, gt, center, ncrops, inv_ts, inv_tt, grad_out=1.0):
    """torch-CPU emulation of dino_head_forward / dino_head_backward: same operands (bf16 where the kernels round),
    same algebra (centre as three bf16 K columns, U = Q.W_s, dX from O and U, dW from (P_s;Q)^T(cnt x_hat; -X~),
    weight-norm Jacobian), fp32 accumulation.  Returns (loss, colsum, dx, dv, dg)."""
    import torch
    xs, xt, vs, vt = xs.float(), xt.float(), vs.float(), vt.float()
    gs, gt, c = gs.float().reshape(-1), gt.float().reshape(-1), center.float().reshape(-1)
    B = xt.shape[0] // 2
    K, D = vs.shape
    rs = ncrops * B
    # dh_prep_rows
    inv_xs = 1.0 / xs.norm(dim=1).clamp_min(1e-12)
    xs_hat = _bf16(xs * inv_xs[:, None])
    xt_hat = _bf16(xt / xt.norm(dim=1, keepdim=True).clamp_min(1e-12))
    ones3 = torch.ones(2 * B, 3)
    xt_aug = torch.cat([xt_hat, ones3], 1)
    xsum = xt_hat.sum(0)
    # dh_prep_weight
    inv_w = 1.0 / vs.norm(dim=1)
    ws = _bf16(vs * (gs * inv_w)[:, None])
    wt = _bf16(vt * (gt / vt.norm(dim=1))[:, None])
    hi = _bf16(c); mid = _bf16(c - hi); lo = _bf16((c - hi) - mid)
    wt_aug = torch.cat([wt, -hi[:, None], -mid[:, None], -lo[:, None]], 1)
    colsum = wt @ xsum
    # teacher statistics + Q (bf16) + U
    zt = (xt_aug @ wt_aug.t()) * inv_tt
    lse_t = torch.logsumexp(zt, 1, keepdim=True)
    Q = _bf16(torch.exp(zt - lse_t))
    U = Q @ ws
    # student statistics + loss
    zs = (xs_hat @ ws.t()) * inv_ts
    lse_s = torch.logsumexp(zs, 1)
    n_terms = 2 * ncrops - 2
    total = 0.0
    for v in range(ncrops):
        cnt = 1.0 if v < 2 else 2.0
        rows = slice(v * B, (v + 1) * B)
        u = sum(U[iq * B:(iq + 1) * B] for iq in range(2) if iq != v)
        total = total + (cnt * lse_s[rows]).sum() - inv_ts * (u * xs_hat[rows]).sum()
    loss = total / (n_terms * B)
    # backward
    coef = inv_ts / (n_terms * B) * grad_out
    P = _bf16(torch.exp(zs - lse_s[:, None]))
    O = P @ ws
    dx = torch.empty(rs, D)
    y = torch.empty(rs + 2 * B, D)
    for v in range(ncrops):
        cnt = 1.0 if v < 2 else 2.0
        rows = slice(v * B, (v + 1) * B)
        u = sum(U[iq * B:(iq + 1) * B] for iq in range(2) if iq != v)
        d = coef * (cnt * O[rows] - u)
        dot = (d * xs_hat[rows]).sum(1, keepdim=True)
        dx[rows] = (d - xs_hat[rows] * dot) * inv_xs[rows, None]
        y[rows] = cnt * xs_hat[rows]
    for iq in range(2):
        y[rs + iq * B: rs + (iq + 1) * B] = _bf16(-sum(xs_hat[v * B:(v + 1) * B] for v in range(ncrops) if v != iq))
    dw_raw = torch.cat([P, Q], 0).t() @ y
    v_hat = vs * inv_w[:, None]
    dot = (dw_raw * v_hat).sum(1)
    dv = (coef * gs * inv_w)[:, None] * (dw_raw - v_hat * dot[:, None])
    dg = coef * dot
    return loss, colsum, dx, dv, dg
