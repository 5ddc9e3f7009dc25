// (f1) DINOHead.last_layer fused with DINOLoss: the [(ncrops+2)B, K] logits never exist in HBM.
// Replaces  x = F.normalize(x); x = self.last_layer(x)   (vision_transformer.py:284,296-300; last_layer is a
// weight-normed Linear(bottleneck -> K, bias=False)) followed by DINOLoss.forward / update_center
// (lafs_train.py:643-679).
//
// The contractions run on the margin head's tcgen05 kernels (head.cu / head_bwd.cu); this file holds the
// streaming kernels around them.  What makes the composition possible without a [rows, K] tensor:
//   * centre subtraction inside the GEMM: the teacher operands get 64 extra K columns, features
//     [x_hat | 1 1 1 | 0..] and prototypes [w | -c_hi -c_mid -c_lo | 0..] with c = c_hi + c_mid + c_lo an exact
//     three-term bf16 split of the fp32 centre, so the accumulator holds  <x_hat, w_k> - c_k  (the products of the
//     extra columns are exact in fp32);
//   * sum_k q_k = 1 turns the cross terms into  sum_k q_k s_k = <U, x_hat_s>  with  U = Q . W_s  ([2B, D]), and the
//     gradients into  dX_hat_s = coef (cnt_v P_s W_s - sum_{iq != v} U_iq),  dW_s = coef (cnt P_s ; -Q)^T (X_hat_s ; X~)
//     with X~_iq = sum_{v != iq} x_hat_s[v]: only softmax probabilities P_s, Q (bf16) are written, by the
//     recomputing gradient GEMM, and only in the backward pass (Q: 2B rows in the forward);
//   * the teacher column sums of update_center are linear in the features:
//     sum_rows t[:, k] = <w_k, sum_rows x_hat_t>, a by-product of the prototype preparation pass.
#include <math.h>
#include "common.cuh"
#include "../../include/lafs_b200.h"

namespace lafs {

constexpr int kDhExtra = 64;     // K columns appended to the teacher operands (3 used)
constexpr float kLn2D = 0.6931471805599453f;

__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// rows of x [R, D] -> out [R, ld] bf16: L2-normalised (F.normalize, eps 1e-12); ld > D: three ones then zeros
template <typename T>
__global__ void __launch_bounds__(256)
dh_prep_rows_kernel(const T* __restrict__ x, int R, int D, int ld, __nv_bfloat16* __restrict__ out,
                    float* __restrict__ inv_norm) {
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= R) return;
  const T* src = x + (size_t)r * D;
  float ss = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float v = (float)src[i];
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  if (lane == 0 && inv_norm != nullptr) inv_norm[r] = inv;
  __nv_bfloat16* dst = out + (size_t)r * ld;
  for (int i = lane; i < D; i += 32) dst[i] = __float2bfloat16_rn((float)src[i] * inv);
  for (int i = D + lane; i < ld; i += 32) dst[i] = __float2bfloat16_rn(i < D + 3 ? 1.f : 0.f);
}

// xsum[d] = sum_r x_hat[r, d]  (fp32, fixed order: bit-reproducible).  One CTA; warp = row group (rows w, w+32, ...),
// lane = 8 consecutive columns (one 16-byte load per row) of the current 256-column chunk.  D % 8 == 0.
__global__ void __launch_bounds__(1024)
dh_xsum_kernel(const __nv_bfloat16* __restrict__ x_hat, int R, int D, int ld, float* __restrict__ xsum) {
  pdl_wait();
  __shared__ float part[32][257];
  const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5;
  for (int c0 = 0; c0 < D; c0 += 256) {
    const int d = c0 + lane * 8;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (d < D) {
#pragma unroll 4
      for (int r = rg; r < R; r += 32) {
        const uint4 q = *reinterpret_cast<const uint4*>(x_hat + (size_t)r * ld + d);
        acc[0] += Half2Ops<__nv_bfloat16>::lo(q.x); acc[1] += Half2Ops<__nv_bfloat16>::hi(q.x);
        acc[2] += Half2Ops<__nv_bfloat16>::lo(q.y); acc[3] += Half2Ops<__nv_bfloat16>::hi(q.y);
        acc[4] += Half2Ops<__nv_bfloat16>::lo(q.z); acc[5] += Half2Ops<__nv_bfloat16>::hi(q.z);
        acc[6] += Half2Ops<__nv_bfloat16>::lo(q.w); acc[7] += Half2Ops<__nv_bfloat16>::hi(q.w);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) part[rg][lane * 8 + j] = acc[j];
    __syncthreads();
    if (threadIdx.x < 256 && c0 + threadIdx.x < D) {
      float t = 0.f;
#pragma unroll
      for (int g = 0; g < 32; ++g) t += part[g][threadIdx.x];
      xsum[c0 + threadIdx.x] = t;
    }
    __syncthreads();
  }
}

// prototypes: weight_norm rows  w_k = v_k * (g_k / ||v_k||)  (torch._weight_norm, dim 0) -> bf16 [K, ld];
// ld > D: columns D..D+2 = -(three-term bf16 split of center[k]), zeros after.  inv_norm[k] = 1/||v_k||.
// colsum[k] = <bf16(w_k), xsum>: the column sum of the teacher logits the tensor cores will see.
template <int NJ>
__global__ void __launch_bounds__(256)
dh_prep_weight_kernel(const float* __restrict__ v, const float* __restrict__ g, const float* __restrict__ center,
                      const float* __restrict__ xsum, int K, int D, int ld, __nv_bfloat16* __restrict__ out,
                      float* __restrict__ inv_norm, float* __restrict__ colsum) {
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = gridDim.x * 8;
  int k = blockIdx.x * 8 + warp;
  if (k >= K) return;
  // a lane owns NJ float4 of a row (D <= 128 NJ, D % 8 == 0); the row is read once, and the next row of this warp is
  // in flight while the current one is reduced, converted and stored
  float4 cur[NJ], nxt[NJ];
  float4 xs[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int d = j * 128 + lane * 4;
    xs[j] = (xsum != nullptr && d < D) ? *reinterpret_cast<const float4*>(xsum + d) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (d < D) cur[j] = ld_stream_f4(v + (size_t)k * D + d);
  }
  for (; k < K; k += nwarps) {
    const int kn = k + nwarps;
    if (kn < K) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int d = j * 128 + lane * 4;
        if (d < D) nxt[j] = ld_stream_f4(v + (size_t)kn * D + d);
      }
    }
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      if (j * 128 + lane * 4 < D) {
        ss = fmaf(cur[j].x, cur[j].x, ss); ss = fmaf(cur[j].y, cur[j].y, ss);
        ss = fmaf(cur[j].z, cur[j].z, ss); ss = fmaf(cur[j].w, cur[j].w, ss);
      }
    }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    const float scale = (g != nullptr ? g[k] : 1.f) / nrm;   // torch: v * (g / norm(v)); a zero row gives inf/nan there too
    if (lane == 0 && inv_norm != nullptr) inv_norm[k] = 1.f / nrm;
    __nv_bfloat16* dst = out + (size_t)k * ld;
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int d = j * 128 + lane * 4;
      if (d < D) {
        uint2 o;
        o.x = Half2Ops<__nv_bfloat16>::pack(cur[j].x * scale, cur[j].y * scale);
        o.y = Half2Ops<__nv_bfloat16>::pack(cur[j].z * scale, cur[j].w * scale);
        *reinterpret_cast<uint2*>(dst + d) = o;
        dot = fmaf(Half2Ops<__nv_bfloat16>::lo(o.x), xs[j].x, dot); dot = fmaf(Half2Ops<__nv_bfloat16>::hi(o.x), xs[j].y, dot);
        dot = fmaf(Half2Ops<__nv_bfloat16>::lo(o.y), xs[j].z, dot); dot = fmaf(Half2Ops<__nv_bfloat16>::hi(o.y), xs[j].w, dot);
      }
    }
    if (colsum != nullptr) {
      dot = warp_sum(dot);
      if (lane == 0) colsum[k] = dot;
    }
    if (ld > D) {
      // 64 extra columns = 128 bytes: lane L writes columns D+2L, D+2L+1 as one 32-bit store
      float e0 = 0.f, e1 = 0.f;
      if (lane < 2) {
        const float c = center != nullptr ? center[k] : 0.f;
        const float hi = bf16_round(c);
        const float mid = bf16_round(c - hi);
        const float lo = bf16_round((c - hi) - mid);
        e0 = lane == 0 ? -hi : -lo;
        e1 = lane == 0 ? -mid : 0.f;
      }
      *reinterpret_cast<uint32_t*>(dst + D + 2 * lane) = Half2Ops<__nv_bfloat16>::pack(e0, e1);
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) cur[j] = nxt[j];
  }
}

// merged head statistics (max2, sum-exp, ., .) -> log2-domain lse of every row
__global__ void dh_lse2_kernel(const float* __restrict__ stats, int R, float* __restrict__ lse2) {
  pdl_wait();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float4 s = *reinterpret_cast<const float4*>(stats + (size_t)r * 4);
  lse2[r] = s.x + log2f(s.y);
}

// loss = 1/((2 ncrops - 2) B) sum_{iq<2} sum_{v != iq} sum_i [ lse(s_v,i / ts) - <U_iq,i , x_hat_v,i> / ts ]
// (SURVEY 8a single-pass identity).  One warp per sample i writes the sample's sum; dh_sum_kernel adds the B sums in
// a fixed order.  A lane owns 8 consecutive columns of every 256-column chunk (D % 8 == 0).
__device__ __forceinline__ void bf16x8_to_float(const uint4 q, float (&f)[8]) {
  f[0] = Half2Ops<__nv_bfloat16>::lo(q.x); f[1] = Half2Ops<__nv_bfloat16>::hi(q.x);
  f[2] = Half2Ops<__nv_bfloat16>::lo(q.y); f[3] = Half2Ops<__nv_bfloat16>::hi(q.y);
  f[4] = Half2Ops<__nv_bfloat16>::lo(q.z); f[5] = Half2Ops<__nv_bfloat16>::hi(q.z);
  f[6] = Half2Ops<__nv_bfloat16>::lo(q.w); f[7] = Half2Ops<__nv_bfloat16>::hi(q.w);
}

__global__ void __launch_bounds__(256)
dh_loss_rows_kernel(const float* __restrict__ lse2_s, const float* __restrict__ U, const __nv_bfloat16* __restrict__ x_hat_s,
                    int B, int ncrops, int D, float inv_ts, float* __restrict__ sample_loss) {
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + warp;
  if (i >= B) return;
  float acc = 0.f;
  for (int d = lane * 8; d < D; d += 256) {
    float u0[8], u1[8];
    {
      const float4 a = *reinterpret_cast<const float4*>(U + (size_t)i * D + d);
      const float4 b = *reinterpret_cast<const float4*>(U + (size_t)i * D + d + 4);
      const float4 c = *reinterpret_cast<const float4*>(U + ((size_t)B + i) * D + d);
      const float4 e = *reinterpret_cast<const float4*>(U + ((size_t)B + i) * D + d + 4);
      u0[0] = a.x; u0[1] = a.y; u0[2] = a.z; u0[3] = a.w; u0[4] = b.x; u0[5] = b.y; u0[6] = b.z; u0[7] = b.w;
      u1[0] = c.x; u1[1] = c.y; u1[2] = c.z; u1[3] = c.w; u1[4] = e.x; u1[5] = e.y; u1[6] = e.z; u1[7] = e.w;
    }
    for (int v = 0; v < ncrops; ++v) {
      float xf[8];
      bf16x8_to_float(*reinterpret_cast<const uint4*>(x_hat_s + ((size_t)v * B + i) * D + d), xf);
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float u = (v != 0 ? u0[j] : 0.f) + (v != 1 ? u1[j] : 0.f);
        dot = fmaf(u, xf[j], dot);
      }
      acc = fmaf(-inv_ts, dot, acc);
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    for (int v = 0; v < ncrops; ++v) acc = fmaf((v < 2 ? 1.f : 2.f) * kLn2D, lse2_s[(size_t)v * B + i], acc);
    sample_loss[i] = acc;
  }
}

// out = scale * sum_i x[i]  (one CTA, fixed order)
__global__ void __launch_bounds__(256)
dh_sum_kernel(const float* __restrict__ x, int n, float scale, float* __restrict__ out) {
  pdl_wait();
  __shared__ float red[256];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) acc += x[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0] * scale;
}

// backward, per row of the [(ncrops+2) B] row space (student rows first, teacher rows after):
//   student row (v,i): d = coef (cnt_v O[v,i] - sum_{iq != v} U[iq,i]),  dx = (d - x_hat <x_hat, d>) inv_norm   (F.normalize
//                      backward), and the dW operand row  y = cnt_v x_hat[v,i]  (bf16, exact);
//   teacher row (iq,i): y = -sum_{v != iq} x_hat[v,i]  (bf16).
// coef = inv_ts / ((2 ncrops - 2) B) * grad_out.
__global__ void __launch_bounds__(256)
dh_bwd_rows_kernel(const float* __restrict__ O, const float* __restrict__ U, const __nv_bfloat16* __restrict__ x_hat_s,
                   const float* __restrict__ inv_norm_s, const float* __restrict__ grad_out, int B, int ncrops, int D,
                   float coef, float* __restrict__ dx, __nv_bfloat16* __restrict__ y) {
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  const int ns = ncrops * B;
  if (row >= ns + 2 * B) return;
  if (row >= ns) {
    const int iq = (row - ns) / B, i = (row - ns) - iq * B;
    for (int d = lane * 8; d < D; d += 256) {
      float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int v = 0; v < ncrops; ++v) {
        if (v == iq) continue;
        float xf[8];
        bf16x8_to_float(*reinterpret_cast<const uint4*>(x_hat_s + ((size_t)v * B + i) * D + d), xf);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += xf[j];
      }
      uint4 o;
      o.x = Half2Ops<__nv_bfloat16>::pack(-acc[0], -acc[1]); o.y = Half2Ops<__nv_bfloat16>::pack(-acc[2], -acc[3]);
      o.z = Half2Ops<__nv_bfloat16>::pack(-acc[4], -acc[5]); o.w = Half2Ops<__nv_bfloat16>::pack(-acc[6], -acc[7]);
      *reinterpret_cast<uint4*>(y + (size_t)row * D + d) = o;
    }
    return;
  }
  const int v = row / B, i = row - v * B;
  const float cnt = v < 2 ? 1.f : 2.f;
  const float cf = coef * __ldg(grad_out);
  // D <= 768: a lane holds its <= 3 groups of 8 columns of d and x_hat in registers between the two passes
  float dd[3][8], xf[3][8];
  float dot = 0.f;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int d = q * 256 + lane * 8;
    if (d < D) {
      bf16x8_to_float(*reinterpret_cast<const uint4*>(x_hat_s + (size_t)row * D + d), xf[q]);
      float o[8], u[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      {
        const float4 a = *reinterpret_cast<const float4*>(O + (size_t)row * D + d);
        const float4 b = *reinterpret_cast<const float4*>(O + (size_t)row * D + d + 4);
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
      }
#pragma unroll
      for (int iq = 0; iq < 2; ++iq) {
        if (iq == v) continue;
        const float4 a = *reinterpret_cast<const float4*>(U + ((size_t)iq * B + i) * D + d);
        const float4 b = *reinterpret_cast<const float4*>(U + ((size_t)iq * B + i) * D + d + 4);
        u[0] += a.x; u[1] += a.y; u[2] += a.z; u[3] += a.w; u[4] += b.x; u[5] += b.y; u[6] += b.z; u[7] += b.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dd[q][j] = cf * (cnt * o[j] - u[j]);
        dot = fmaf(dd[q][j], xf[q][j], dot);
      }
    }
  }
  dot = warp_sum(dot);
  const float inv = inv_norm_s[row];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int d = q * 256 + lane * 8;
    if (d < D) {
      float r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = (dd[q][j] - xf[q][j] * dot) * inv;
      *reinterpret_cast<float4*>(dx + (size_t)row * D + d) = make_float4(r[0], r[1], r[2], r[3]);
      *reinterpret_cast<float4*>(dx + (size_t)row * D + d + 4) = make_float4(r[4], r[5], r[6], r[7]);
      uint4 o;
      o.x = Half2Ops<__nv_bfloat16>::pack(cnt * xf[q][0], cnt * xf[q][1]); o.y = Half2Ops<__nv_bfloat16>::pack(cnt * xf[q][2], cnt * xf[q][3]);
      o.z = Half2Ops<__nv_bfloat16>::pack(cnt * xf[q][4], cnt * xf[q][5]); o.w = Half2Ops<__nv_bfloat16>::pack(cnt * xf[q][6], cnt * xf[q][7]);
      *reinterpret_cast<uint4*>(y + (size_t)row * D + d) = o;
    }
  }
}

// weight-norm backward of  w = g v / ||v||  from the raw (unscaled) dW [K, D]:
//   dv = coef g/||v|| (dW - v_hat <v_hat, dW>),   dg = coef <v_hat, dW>,   v_hat = v / ||v||.   dv may alias dw.
__global__ void __launch_bounds__(256)
dh_wn_bwd_kernel(const float* dw, const float* __restrict__ v, const float* __restrict__ g,
                 const float* __restrict__ inv_norm, const float* __restrict__ grad_out, int K, int D, float coef,
                 float* dv, float* __restrict__ dg) {
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k = blockIdx.x * 8 + warp;
  if (k >= K) return;
  const float inv = inv_norm[k];
  const float* dr = dw + (size_t)k * D;
  const float* vr = v + (size_t)k * D;
  float dot = 0.f;
  for (int d = lane * 4; d < D; d += 128) {
    const float4 a = *reinterpret_cast<const float4*>(dr + d);
    const float4 b = *reinterpret_cast<const float4*>(vr + d);
    dot = fmaf(a.x, b.x * inv, dot); dot = fmaf(a.y, b.y * inv, dot);
    dot = fmaf(a.z, b.z * inv, dot); dot = fmaf(a.w, b.w * inv, dot);
  }
  dot = warp_sum(dot);
  const float cf = coef * __ldg(grad_out);
  const float sc = cf * (g != nullptr ? g[k] : 1.f) * inv;
  if (lane == 0 && dg != nullptr) dg[k] = cf * dot;
  for (int d = lane * 4; d < D; d += 128) {
    const float4 a = *reinterpret_cast<const float4*>(dr + d);
    const float4 b = *reinterpret_cast<const float4*>(vr + d);
    float4 o;
    o.x = sc * (a.x - b.x * inv * dot); o.y = sc * (a.y - b.y * inv * dot);
    o.z = sc * (a.z - b.z * inv * dot); o.w = sc * (a.w - b.w * inv * dot);
    *reinterpret_cast<float4*>(dv + (size_t)k * D + d) = o;
  }
}

}  // namespace lafs

using namespace lafs;

extern "C" int lafs_dh_extra_cols(void) { return kDhExtra; }

extern "C" int lafs_dh_prep_rows(const void* x, int dtype, int R, int D, int ld, void* out_bf16, float* inv_norm,
                                 lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(x)) return brc;
  LAFS_REQUIRE(x && out_bf16 && R >= 0 && D > 0, LAFS_ERR_ARG, "lafs_dh_prep_rows: bad argument");
  LAFS_REQUIRE(ld == D || ld == D + kDhExtra, LAFS_ERR_ARG, "lafs_dh_prep_rows: ld=%d must be D or D+%d", ld, kDhExtra);
  LAFS_REQUIRE(dtype >= 0 && dtype <= 2, LAFS_ERR_ARG, "lafs_dh_prep_rows: dtype=%d", dtype);
  if (R == 0) return LAFS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (R + 7) / 8;
  __nv_bfloat16* o = (__nv_bfloat16*)out_bf16;
  if (dtype == LAFS_F32) launch_pdl((dh_prep_rows_kernel<float>), dim3(grid), dim3(256), (size_t)0, st, (const float*)x, R, D, ld, o, inv_norm);
  else if (dtype == LAFS_BF16) launch_pdl((dh_prep_rows_kernel<__nv_bfloat16>), dim3(grid), dim3(256), (size_t)0, st, (const __nv_bfloat16*)x, R, D, ld, o, inv_norm);
  else launch_pdl((dh_prep_rows_kernel<__half>), dim3(grid), dim3(256), (size_t)0, st, (const __half*)x, R, D, ld, o, inv_norm);
  return check_launch("lafs_dh_prep_rows");
}

extern "C" int lafs_dh_xsum(const void* x_hat_bf16, int R, int D, int ld, float* xsum, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(x_hat_bf16)) return brc;
  LAFS_REQUIRE(x_hat_bf16 && xsum && R > 0 && D > 0 && D % 8 == 0 && ld >= D && ld % 8 == 0, LAFS_ERR_ARG,
               "lafs_dh_xsum: bad argument (D and ld multiples of 8)");
  LAFS_REQUIRE(((uintptr_t)x_hat_bf16 & 15u) == 0, LAFS_ERR_ARG, "lafs_dh_xsum: x_hat must be 16-byte aligned");
  launch_pdl((dh_xsum_kernel), dim3(1), dim3(1024), (size_t)0, (cudaStream_t)stream, (const __nv_bfloat16*)x_hat_bf16, R, D, ld, xsum);
  return check_launch("lafs_dh_xsum");
}

extern "C" int lafs_dh_prep_weight(const float* weight_v, const float* weight_g, const float* center, const float* xsum,
                                   int K, int D, int ld, void* out_bf16, float* inv_norm, float* colsum,
                                   lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(weight_v)) return brc;
  LAFS_REQUIRE(weight_v && out_bf16 && K > 0 && D > 0, LAFS_ERR_ARG, "lafs_dh_prep_weight: bad argument");
  LAFS_REQUIRE(D % 8 == 0 && D <= 1024, LAFS_ERR_ARG, "lafs_dh_prep_weight: D=%d must be a multiple of 8, <= 1024", D);
  LAFS_REQUIRE((((uintptr_t)weight_v | (uintptr_t)xsum) & 15u) == 0 && ((uintptr_t)out_bf16 & 7u) == 0, LAFS_ERR_ARG,
               "lafs_dh_prep_weight: weight_v / xsum must be 16-byte aligned");
  LAFS_REQUIRE(ld == D || ld == D + kDhExtra, LAFS_ERR_ARG, "lafs_dh_prep_weight: ld=%d must be D or D+%d", ld, kDhExtra);
  LAFS_REQUIRE(!(center != nullptr && ld == D), LAFS_ERR_ARG, "lafs_dh_prep_weight: a centre needs ld = D+%d", kDhExtra);
  LAFS_REQUIRE((colsum == nullptr) == (xsum == nullptr), LAFS_ERR_ARG, "lafs_dh_prep_weight: colsum and xsum go together");
  // persistent warps: up to 8 CTAs of 8 warps per SM, every warp walks rows with the grid stride
  int grid = (K + 7) / 8;
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* o = (__nv_bfloat16*)out_bf16;
  const int nj = (D + 127) / 128;
  if (nj <= 1) launch_pdl((dh_prep_weight_kernel<1>), dim3(grid), dim3(256), (size_t)0, st, weight_v, weight_g, center, xsum, K, D, ld, o, inv_norm, colsum);
  else if (nj == 2) launch_pdl((dh_prep_weight_kernel<2>), dim3(grid), dim3(256), (size_t)0, st, weight_v, weight_g, center, xsum, K, D, ld, o, inv_norm, colsum);
  else if (nj <= 4) launch_pdl((dh_prep_weight_kernel<4>), dim3(grid), dim3(256), (size_t)0, st, weight_v, weight_g, center, xsum, K, D, ld, o, inv_norm, colsum);
  else launch_pdl((dh_prep_weight_kernel<8>), dim3(grid), dim3(256), (size_t)0, st, weight_v, weight_g, center, xsum, K, D, ld, o, inv_norm, colsum);
  return check_launch("lafs_dh_prep_weight");
}

extern "C" int lafs_dh_lse2(const float* row_stats, int R, float* lse2, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(row_stats)) return brc;
  LAFS_REQUIRE(row_stats && lse2 && R > 0, LAFS_ERR_ARG, "lafs_dh_lse2: bad argument");
  launch_pdl((dh_lse2_kernel), dim3((R + 255) / 256), dim3(256), (size_t)0, (cudaStream_t)stream, row_stats, R, lse2);
  return check_launch("lafs_dh_lse2");
}

extern "C" int lafs_dh_loss(const float* lse2_s, const float* U, const void* x_hat_s_bf16, int B, int ncrops, int D,
                            float inv_student_temp, float* sample_loss, float* loss_out, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(lse2_s)) return brc;
  LAFS_REQUIRE(lse2_s && U && x_hat_s_bf16 && sample_loss && loss_out, LAFS_ERR_ARG, "lafs_dh_loss: null pointer");
  LAFS_REQUIRE(B > 0 && ncrops >= 2 && D > 0 && D % 8 == 0, LAFS_ERR_ARG, "lafs_dh_loss: B=%d ncrops=%d D=%d (D %% 8 == 0)", B, ncrops, D);
  LAFS_REQUIRE((((uintptr_t)U | (uintptr_t)x_hat_s_bf16) & 15u) == 0, LAFS_ERR_ARG, "lafs_dh_loss: U / x_hat_s must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  launch_pdl((dh_loss_rows_kernel), dim3((B + 7) / 8), dim3(256), (size_t)0, st, lse2_s, U, (const __nv_bfloat16*)x_hat_s_bf16,
             B, ncrops, D, inv_student_temp, sample_loss);
  launch_pdl((dh_sum_kernel), dim3(1), dim3(256), (size_t)0, st, (const float*)sample_loss, B, 1.f / ((float)(2 * ncrops - 2) * (float)B), loss_out);
  return check_launch("lafs_dh_loss");
}

extern "C" int lafs_dh_bwd_rows(const float* O, const float* U, const void* x_hat_s_bf16, const float* inv_norm_s,
                                const float* grad_out, int B, int ncrops, int D, float inv_student_temp, float* dx,
                                void* y_bf16, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(O)) return brc;
  LAFS_REQUIRE(O && U && x_hat_s_bf16 && inv_norm_s && grad_out && dx && y_bf16, LAFS_ERR_ARG, "lafs_dh_bwd_rows: null pointer");
  LAFS_REQUIRE(B > 0 && ncrops >= 2 && D > 0 && D % 8 == 0 && D <= 768, LAFS_ERR_ARG, "lafs_dh_bwd_rows: B=%d ncrops=%d D=%d (D %% 8 == 0, <= 768)", B, ncrops, D);
  LAFS_REQUIRE((((uintptr_t)O | (uintptr_t)U | (uintptr_t)x_hat_s_bf16 | (uintptr_t)dx | (uintptr_t)y_bf16) & 15u) == 0, LAFS_ERR_ARG,
               "lafs_dh_bwd_rows: pointers must be 16-byte aligned");
  const float coef = inv_student_temp / ((float)(2 * ncrops - 2) * (float)B);
  const int rows = (ncrops + 2) * B;
  launch_pdl((dh_bwd_rows_kernel), dim3((rows + 7) / 8), dim3(256), (size_t)0, (cudaStream_t)stream, O, U,
             (const __nv_bfloat16*)x_hat_s_bf16, inv_norm_s, grad_out, B, ncrops, D, coef, dx, (__nv_bfloat16*)y_bf16);
  return check_launch("lafs_dh_bwd_rows");
}

extern "C" int lafs_dh_wn_bwd(const float* dw_raw, const float* weight_v, const float* weight_g, const float* inv_norm,
                              const float* grad_out, int K, int D, float coef, float* grad_v, float* grad_g,
                              lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(dw_raw)) return brc;
  LAFS_REQUIRE(dw_raw && weight_v && inv_norm && grad_out && grad_v, LAFS_ERR_ARG, "lafs_dh_wn_bwd: null pointer");
  LAFS_REQUIRE(K > 0 && D > 0 && D % 4 == 0, LAFS_ERR_ARG, "lafs_dh_wn_bwd: K=%d D=%d (D must be a multiple of 4)", K, D);
  LAFS_REQUIRE((((uintptr_t)dw_raw | (uintptr_t)weight_v | (uintptr_t)grad_v) & 15u) == 0, LAFS_ERR_ARG,
               "lafs_dh_wn_bwd: pointers must be 16-byte aligned");
  launch_pdl((dh_wn_bwd_kernel), dim3((K + 7) / 8), dim3(256), (size_t)0, (cudaStream_t)stream, dw_raw, weight_v, weight_g, inv_norm,
             grad_out, K, D, coef, grad_v, grad_g);
  return check_launch("lafs_dh_wn_bwd");
}
