// (1a) Landmark-guided bilinear patch gather (stand-alone fp32 form) and landmark
// post-processing.  Replaces extract_patches_pytorch_gridsample
// (face_pre_pro/ViT_face.py:1615-1656: n python iterations x (3 element-wise + 1 grid_sample
// launch) + stack/permute copies) by one launch, and the ~10 tiny kernels of the landmark
// tail (ViT_face.py:1347-1378) by one launch.
//
// fp32 op order per sample point is the reference's (tests/kernel_emulation.py shows it is
// bit-identical on CPU):  p = theta + off;  g = p/(H/2) - 1;  ix = fma(g + 1, H/2, -0.5);
// weights (1-w)(1-n) etc.;  out = fma(v_se,se, fma(v_sw,sw, fma(v_ne,ne, v_nw*nw))).
#include "common.cuh"
#include "../../include/lafs_b200.h"

namespace lafs {

constexpr int kPatch = 8;
constexpr int kPatchPix = kPatch * kPatch;   // 64 sample points per patch
constexpr int kGatherThreads = 256;
constexpr int kPatchesPerCta = kGatherThreads / kPatchPix;  // 4
constexpr int kMaxC = 4;

// Source coordinate of sample index `idx` (0..7) around landmark coordinate `theta`.
// RECIP=false: IEEE division (reference CPU path); RECIP=true: multiply by fp32(1/half)
// (what eager CUDA does for tensor / python_scalar).
// The reference normalises BOTH coordinates with imgs.shape[2]*0.5 (ViT_face.py:1630,1645);
// grid_sample then un-normalises x with W/2 and y with H/2.
template <bool RECIP>
__device__ __forceinline__ float source_coord(float theta, int idx, float norm_half, float unnorm_half) {
  const float p = __fadd_rn(theta, (float)(idx - kPatch / 2));
  const float q = RECIP ? __fmul_rn(p, 1.0f / norm_half) : __fdiv_rn(p, norm_half);
  const float g = __fsub_rn(q, 1.0f);
  const float a = __fadd_rn(g, 1.0f);
  return __fmaf_rn(a, unnorm_half, -0.5f);
}

struct Bilinear {
  int x0, y0;
  float nw, ne, sw, se;  // weights
  float e, w, s, n;      // 1-fx, fx, 1-fy, fy
};

template <bool RECIP>
__device__ __forceinline__ Bilinear make_bilinear(float tx, float ty, int i, int j, int H, int W) {
  Bilinear bl;
  const float ix = source_coord<RECIP>(tx, i, 0.5f * (float)H, 0.5f * (float)W);
  const float iy = source_coord<RECIP>(ty, j, 0.5f * (float)H, 0.5f * (float)H);
  const float fx = floorf(ix), fy = floorf(iy);
  bl.w = __fsub_rn(ix, fx); bl.e = __fsub_rn(1.0f, bl.w);
  bl.n = __fsub_rn(iy, fy); bl.s = __fsub_rn(1.0f, bl.n);
  bl.nw = __fmul_rn(bl.s, bl.e); bl.ne = __fmul_rn(bl.s, bl.w);
  bl.sw = __fmul_rn(bl.n, bl.e); bl.se = __fmul_rn(bl.n, bl.w);
  // clamp before the int conversion: landmarks may be arbitrarily far outside the image
  bl.x0 = (int)fminf(fmaxf(fx, -2.0f), (float)W + 1.0f);
  bl.y0 = (int)fminf(fmaxf(fy, -2.0f), (float)H + 1.0f);
  return bl;
}

__device__ __forceinline__ float pix(const float* __restrict__ plane, int x, int y, int H, int W) {
  return (x >= 0 && x < W && y >= 0 && y < H) ? __ldg(plane + y * W + x) : 0.0f;
}

// thread t of a patch: i = t & 7 steps x (image columns, contiguous in memory -> coalesced
// reads), j = t >> 3 steps y.  Output patch pixel (i, j) = out row i, column j.
template <bool RECIP>
__global__ void __launch_bounds__(kGatherThreads)
gather_fwd_kernel(const float* __restrict__ imgs, const float* __restrict__ theta, float* __restrict__ out,
                  int total_patches, int C, int H, int W, int n, int r, int layout) {
  pdl_wait();
  __shared__ float tile[kPatchesPerCta][kPatchPix * kMaxC];
  const int p_local = threadIdx.x >> 6, t = threadIdx.x & 63;
  const int i = t & 7, j = t >> 3;
  const int patch = blockIdx.x * kPatchesPerCta + p_local;
  if (patch < total_patches) {
    const int b = patch / n;
    const float tx = __ldg(theta + 2 * (size_t)patch), ty = __ldg(theta + 2 * (size_t)patch + 1);
    const Bilinear bl = make_bilinear<RECIP>(tx, ty, i, j, H, W);
    const float* img = imgs + (size_t)b * C * H * W;
    for (int c = 0; c < C; ++c) {
      const float* plane = img + (size_t)c * H * W;
      const float vnw = pix(plane, bl.x0, bl.y0, H, W), vne = pix(plane, bl.x0 + 1, bl.y0, H, W);
      const float vsw = pix(plane, bl.x0, bl.y0 + 1, H, W), vse = pix(plane, bl.x0 + 1, bl.y0 + 1, H, W);
      float acc = __fmul_rn(vnw, bl.nw);
      acc = __fmaf_rn(vne, bl.ne, acc);
      acc = __fmaf_rn(vsw, bl.sw, acc);
      acc = __fmaf_rn(vse, bl.se, acc);
      tile[p_local][(i * kPatch + j) * C + c] = acc;
    }
  }
  __syncthreads();
  if (patch >= total_patches) return;
  const int b = patch / n, k = patch - b * n;
  if (layout == LAFS_LAYOUT_TOKENS) {
    float* dst = out + (size_t)patch * kPatchPix * C;
    for (int e = t; e < kPatchPix * C; e += kPatchPix) dst[e] = tile[p_local][e];
  } else {
    const int rr = k / r, qq = k - rr * r;
    const int side = r * kPatch;
    for (int e = t; e < kPatchPix * C; e += kPatchPix) {
      const int c = e >> 6, ii = (e >> 3) & 7, jj = e & 7;
      out[(((size_t)b * C + c) * side + rr * kPatch + ii) * side + qq * kPatch + jj] =
          tile[p_local][(ii * kPatch + jj) * C + c];
    }
  }
}

// Backward: grad_theta (x,y) per landmark and scatter-add into grad_imgs.
template <bool RECIP>
__global__ void __launch_bounds__(kGatherThreads)
gather_bwd_kernel(const float* __restrict__ imgs, const float* __restrict__ theta,
                  const float* __restrict__ gout, float* __restrict__ gimgs, float* __restrict__ gtheta,
                  int total_patches, int C, int H, int W, int n, int r, int layout) {
  pdl_wait();
  __shared__ float red[kPatchesPerCta][2][2];
  const int p_local = threadIdx.x >> 6, t = threadIdx.x & 63;
  const int i = t & 7, j = t >> 3;
  const int patch = blockIdx.x * kPatchesPerCta + p_local;
  float gx = 0.f, gy = 0.f;
  if (patch < total_patches) {
    const int b = patch / n, k = patch - b * n;
    const int rr = k / r, qq = k - rr * r, side = r * kPatch;
    const float tx = __ldg(theta + 2 * (size_t)patch), ty = __ldg(theta + 2 * (size_t)patch + 1);
    const Bilinear bl = make_bilinear<RECIP>(tx, ty, i, j, H, W);
    const bool x0_in = bl.x0 >= 0 && bl.x0 < W, x1_in = bl.x0 + 1 >= 0 && bl.x0 + 1 < W;
    const bool y0_in = bl.y0 >= 0 && bl.y0 < H, y1_in = bl.y0 + 1 >= 0 && bl.y0 + 1 < H;
    for (int c = 0; c < C; ++c) {
      float g;
      if (layout == LAFS_LAYOUT_TOKENS) g = __ldg(gout + (size_t)patch * kPatchPix * C + (i * kPatch + j) * C + c);
      else g = __ldg(gout + (((size_t)b * C + c) * side + rr * kPatch + i) * side + qq * kPatch + j);
      const size_t plane_off = ((size_t)b * C + c) * H * W;
      const float* plane = imgs + plane_off;
      const float vnw = pix(plane, bl.x0, bl.y0, H, W), vne = pix(plane, bl.x0 + 1, bl.y0, H, W);
      const float vsw = pix(plane, bl.x0, bl.y0 + 1, H, W), vse = pix(plane, bl.x0 + 1, bl.y0 + 1, H, W);
      gx += g * ((vne - vnw) * bl.s + (vse - vsw) * bl.n);
      gy += g * ((vsw - vnw) * bl.e + (vse - vne) * bl.w);
      if (gimgs != nullptr) {
        float* gp = gimgs + plane_off;
        if (x0_in && y0_in) atomicAdd(gp + bl.y0 * W + bl.x0, g * bl.nw);
        if (x1_in && y0_in) atomicAdd(gp + bl.y0 * W + bl.x0 + 1, g * bl.ne);
        if (x0_in && y1_in) atomicAdd(gp + (bl.y0 + 1) * W + bl.x0, g * bl.sw);
        if (x1_in && y1_in) atomicAdd(gp + (bl.y0 + 1) * W + bl.x0 + 1, g * bl.se);
      }
    }
  }
  if (gtheta == nullptr) return;
  gx = warp_sum(gx);
  gy = warp_sum(gy);
  const int warp_in_patch = (threadIdx.x >> 5) & 1;
  if ((threadIdx.x & 31) == 0) { red[p_local][warp_in_patch][0] = gx; red[p_local][warp_in_patch][1] = gy; }
  __syncthreads();
  if (t < 2 && patch < total_patches) gtheta[2 * (size_t)patch + t] = red[p_local][0][t] + red[p_local][1][t];
}

// ---- landmark tail: joint min-max scaling (+noise, +re-sampling) ------------------------------
// one warp per sample; all of a sample's values are fetched with independent loads first (the
// kernel is pure latency: 2n <= 32*kLmVals values per sample), then reduced and rescaled.
constexpr int kLmVals = 13;   // 13 * 32 = 416 >= 392 = 2 * 196
__global__ void __launch_bounds__(128)
landmark_post_kernel(const float* __restrict__ raw, const float* __restrict__ noise,
                     const int64_t* __restrict__ extract_id, float* __restrict__ theta_out,
                     float* __restrict__ minmax_out, int B, int n, int keep, float scale) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= B) return;
  const int m = 2 * n;
  const float* x = raw + (size_t)b * m;
  const bool small = m <= 32 * kLmVals;
  float v[kLmVals];
  float mn = INFINITY, mx = -INFINITY;
  if (small) {
#pragma unroll
    for (int i = 0; i < kLmVals; ++i) {
      const int e = lane + 32 * i;
      v[i] = e < m ? __ldg(x + e) : NAN;
    }
#pragma unroll
    for (int i = 0; i < kLmVals; ++i) {       // fminf / fmaxf ignore the NaN padding
      mn = fminf(mn, v[i]);
      mx = fmaxf(mx, v[i]);
    }
  } else {
    for (int e = lane; e < m; e += 32) {
      const float t = __ldg(x + e);
      mn = fminf(mn, t);
      mx = fmaxf(mx, t);
    }
  }
  mx = warp_max(mx);
  mn = -warp_max(-mn);
  const float range = __fsub_rn(mx, mn);
  if (minmax_out != nullptr && lane == 0) { minmax_out[2 * b] = mn; minmax_out[2 * b + 1] = mx; }
  // (theta - t_min)/(t_max - t_min)*111, then + noise   (ViT_face.py:1351,1362)
  if (extract_id == nullptr && small) {
#pragma unroll
    for (int i = 0; i < kLmVals; ++i) {
      const int e = lane + 32 * i;
      if (e < m) {
        float t = __fmul_rn(__fdiv_rn(__fsub_rn(v[i], mn), range), scale);
        if (noise != nullptr) t = __fadd_rn(t, __ldg(noise + (size_t)b * m + e));
        theta_out[(size_t)b * m + e] = t;
      }
    }
    return;
  }
  const int n_out = extract_id != nullptr ? keep : n;
  for (int e = lane; e < 2 * n_out; e += 32) {
    const int kk = e >> 1, d = e & 1;
    const int src = extract_id != nullptr ? (int)__ldg(extract_id + (size_t)b * keep + kk) : kk;
    float t = __fmul_rn(__fdiv_rn(__fsub_rn(__ldg(x + 2 * src + d), mn), range), scale);
    if (noise != nullptr) t = __fadd_rn(t, __ldg(noise + ((size_t)b * n + src) * 2 + d));
    theta_out[((size_t)b * n_out + kk) * 2 + d] = t;
  }
}

__global__ void __launch_bounds__(128)
landmark_post_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ gtheta,
                         float* __restrict__ graw, int B, int n, float scale) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= B) return;
  const int m = 2 * n;
  const float* x = raw + (size_t)b * m;
  const float* g = gtheta + (size_t)b * m;
  float mn = INFINITY, mx = -INFINITY;
  int imn = 0x7fffffff, imx = 0x7fffffff;
  for (int e = lane; e < m; e += 32) {
    const float v = __ldg(x + e);
    if (v < mn) { mn = v; imn = e; }
    if (v > mx) { mx = v; imx = e; }
  }
  // arg-reduce, first index wins on ties
  for (int o = 16; o > 0; o >>= 1) {
    const float omn = __shfl_xor_sync(0xffffffffu, mn, o), omx = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oimn = __shfl_xor_sync(0xffffffffu, imn, o), oimx = __shfl_xor_sync(0xffffffffu, imx, o);
    if (omn < mn || (omn == mn && oimn < imn)) { mn = omn; imn = oimn; }
    if (omx > mx || (omx == mx && oimx < imx)) { mx = omx; imx = oimx; }
  }
  const float range = mx - mn;
  const float inv = scale / range;
  float d_mn = 0.f, d_mx = 0.f;
  for (int e = lane; e < m; e += 32) {
    const float ge = __ldg(g + e);
    const float u = (__ldg(x + e) - mn) / range;  // in [0,1]
    d_mn += ge * inv * (u - 1.f);
    d_mx += ge * inv * (-u);
  }
  d_mn = warp_sum(d_mn);
  d_mx = warp_sum(d_mx);
  for (int e = lane; e < m; e += 32) {
    float v = __ldg(g + e) * inv;
    if (e == imn) v += d_mn;
    if (e == imx) v += d_mx;
    graw[(size_t)b * m + e] = v;
  }
}

}  // namespace lafs

static int isqrt_exact(int n) {
  int r = 0;
  while ((r + 1) * (r + 1) <= n) ++r;
  return r * r == n ? r : -1;
}

static int gather_check(const float* imgs, const float* theta, int Bv, int C, int H, int W, int n, int layout,
                        int coord_mode, const char* who, int* r_out) {
  using namespace lafs;
  LAFS_REQUIRE(imgs && theta, LAFS_ERR_ARG, "%s: null pointer", who);
  LAFS_REQUIRE(Bv > 0 && n > 0 && H > 0 && W > 0, LAFS_ERR_ARG, "%s: bad sizes Bv=%d n=%d H=%d W=%d", who, Bv, n, H, W);
  LAFS_REQUIRE(C >= 1 && C <= kMaxC, LAFS_ERR_ARG, "%s: C=%d outside [1,%d]", who, C, kMaxC);
  LAFS_REQUIRE(layout == LAFS_LAYOUT_MOSAIC || layout == LAFS_LAYOUT_TOKENS, LAFS_ERR_ARG, "%s: layout=%d", who, layout);
  LAFS_REQUIRE(coord_mode == LAFS_COORD_DIV || coord_mode == LAFS_COORD_RECIP, LAFS_ERR_ARG, "%s: coord_mode=%d", who, coord_mode);
  int r = isqrt_exact(n);
  LAFS_REQUIRE(layout == LAFS_LAYOUT_TOKENS || r > 0, LAFS_ERR_ARG, "%s: mosaic layout needs a square landmark count, got %d", who, n);
  LAFS_REQUIRE((long long)Bv * n < (1LL << 31), LAFS_ERR_ARG, "%s: too many patches", who);
  *r_out = r > 0 ? r : 1;
  return LAFS_OK;
}

extern "C" int lafs_gather_fwd(const float* imgs, const float* theta, float* out, int Bv, int C, int H, int W,
                               int n, int layout, int coord_mode, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(imgs)) return brc;
  using namespace lafs;
  if (Bv == 0) return LAFS_OK;   // empty batch: nothing to do (pointers may be null)
  int r;
  int rc = gather_check(imgs, theta, Bv, C, H, W, n, layout, coord_mode, "lafs_gather_fwd", &r);
  if (rc) return rc;
  LAFS_REQUIRE(out != nullptr, LAFS_ERR_ARG, "lafs_gather_fwd: null output");
  const int total = Bv * n;
  if (total == 0) return LAFS_OK;
  const int grid = (total + kPatchesPerCta - 1) / kPatchesPerCta;
  cudaStream_t st = (cudaStream_t)stream;
  if (coord_mode == LAFS_COORD_RECIP)
    launch_pdl((gather_fwd_kernel<true>), dim3(grid), dim3(kGatherThreads), (size_t)(0), st, imgs, theta, out, total, C, H, W, n, r, layout);
  else
    launch_pdl((gather_fwd_kernel<false>), dim3(grid), dim3(kGatherThreads), (size_t)(0), st, imgs, theta, out, total, C, H, W, n, r, layout);
  return check_launch("lafs_gather_fwd");
}

extern "C" int lafs_gather_bwd(const float* imgs, const float* theta, const float* grad_out, float* grad_imgs,
                               float* grad_theta, int Bv, int C, int H, int W, int n, int layout,
                               int coord_mode, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(imgs)) return brc;
  using namespace lafs;
  if (Bv == 0) return LAFS_OK;
  int r;
  int rc = gather_check(imgs, theta, Bv, C, H, W, n, layout, coord_mode, "lafs_gather_bwd", &r);
  if (rc) return rc;
  LAFS_REQUIRE(grad_out != nullptr, LAFS_ERR_ARG, "lafs_gather_bwd: null grad_out");
  const int total = Bv * n;
  if (total == 0 || (grad_imgs == nullptr && grad_theta == nullptr)) return LAFS_OK;
  const int grid = (total + kPatchesPerCta - 1) / kPatchesPerCta;
  cudaStream_t st = (cudaStream_t)stream;
  if (coord_mode == LAFS_COORD_RECIP)
    launch_pdl((gather_bwd_kernel<true>), dim3(grid), dim3(kGatherThreads), (size_t)(0), st, imgs, theta, grad_out, grad_imgs, grad_theta, total, C, H, W, n, r, layout);
  else
    launch_pdl((gather_bwd_kernel<false>), dim3(grid), dim3(kGatherThreads), (size_t)(0), st, imgs, theta, grad_out, grad_imgs, grad_theta, total, C, H, W, n, r, layout);
  return check_launch("lafs_gather_bwd");
}

extern "C" int lafs_landmark_post(const float* raw, const float* noise, const int64_t* extract_id,
                                  float* theta_out, float* minmax_out, int B, int n, int keep, float scale,
                                  lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(raw)) return brc;
  using namespace lafs;
  if (B == 0) return LAFS_OK;
  LAFS_REQUIRE(raw && theta_out, LAFS_ERR_ARG, "lafs_landmark_post: null pointer");
  LAFS_REQUIRE(B >= 0 && n > 0, LAFS_ERR_ARG, "lafs_landmark_post: B=%d n=%d", B, n);
  LAFS_REQUIRE(extract_id == nullptr || keep > 0, LAFS_ERR_ARG, "lafs_landmark_post: keep=%d with extract_id", keep);
  if (B == 0) return LAFS_OK;
  launch_pdl((landmark_post_kernel), dim3((B + 3) / 4), dim3(128), (size_t)(0), (cudaStream_t)stream, raw, noise, extract_id, theta_out, minmax_out,
                                                                     B, n, keep, scale);
  return check_launch("lafs_landmark_post");
}

extern "C" int lafs_landmark_post_bwd(const float* raw, const float* grad_theta, float* grad_raw, int B, int n,
                                      float scale, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(raw)) return brc;
  using namespace lafs;
  LAFS_REQUIRE(raw && grad_theta && grad_raw, LAFS_ERR_ARG, "lafs_landmark_post_bwd: null pointer");
  LAFS_REQUIRE(B >= 0 && n > 0, LAFS_ERR_ARG, "lafs_landmark_post_bwd: B=%d n=%d", B, n);
  if (B == 0) return LAFS_OK;
  launch_pdl((landmark_post_bwd_kernel), dim3((B + 3) / 4), dim3(128), (size_t)(0), (cudaStream_t)stream, raw, grad_theta, grad_raw, B, n, scale);
  return check_launch("lafs_landmark_post_bwd");
}
