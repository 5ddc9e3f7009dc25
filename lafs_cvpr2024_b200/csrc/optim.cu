// (f3) Student update as two multi-tensor launches: per-tensor gradient clipping + AdamW + teacher EMA.
// Replaces, for every parameter tensor of the student (147 for ViT-B + DINO head):
//   utils.clip_gradients            utils.py:132-141   one norm kernel + one .item() host sync + one mul_ PER TENSOR
//   utils.cancel_gradients_last_layer  utils.py:144-149   (tensors flagged "no gradient this step" are skipped)
//   torch.optim.AdamW.step          lafs_train.py:400,601-609   ~10 element-wise passes (foreach) over p, g, m, v
//   the teacher EMA loop            lafs_train.py:610-613
// with
//   pass 1  grad_sumsq_kernel + clip_coef_kernel : per-tensor L2 norms and clip coefficients, on the device
//           (4 B/param read; the norms stay in a device array -- no host sync)
//   pass 2  adamw_ema_kernel : g*coef -> decoupled weight decay -> moments -> parameter -> teacher EMA of the
//           NEW parameter, 36 B/param (read p, g, m, v, k; write p, m, v, k) instead of 28 + 12 + 8.
// Arithmetic follows torch.optim.AdamW's single-tensor formulas (fp32, separately rounded where torch rounds):
//   p *= 1 - lr*wd;  m = m + (g - m)*(1-b1);  v = v*b2 + g*g*(1-b2);
//   p -= step_size * m / (sqrt(v)/bc2_sqrt + eps),   step_size = lr/(1-b1^t), bc2_sqrt = sqrt(1-b2^t)
// The step-dependent scalars live in a small DEVICE array (hyper) so that a captured CUDA graph can be replayed
// with new values (the host refreshes the array before the replay).
#include "common.cuh"
#include "../../include/lafs_b200.h"

namespace lafs {

struct OptChunk {
  float* p;         // student parameter run (in place)
  const float* g;   // its gradient run (nullptr: tensor has no gradient -> only the EMA runs)
  float* m;         // exp_avg
  float* v;         // exp_avg_sq
  float* k;         // teacher parameter run (nullptr: no EMA for this tensor)
  int32_t n;        // <= LAFS_EMA_CHUNK
  int32_t tensor;   // index into the per-tensor arrays
};
static_assert(sizeof(OptChunk) == 48, "table record is 48 bytes");

// hyper[] (the host evaluates the scalar expressions in double, like torch's Python scalars, then rounds to fp32):
//   0 decay = 1 - lr*weight_decay (regularised group), 1 step_size = lr/(1-b1^t), 2 bc2_sqrt = sqrt(1-b2^t), 3 eps,
//   4 1-beta1, 5 beta2, 6 1-beta2, 7 ema momentum, 8 1 - ema momentum, 9 clip (<= 0: no clipping)
enum { H_DECAY = 0, H_STEP_SIZE, H_BC2_SQRT, H_EPS, H_OMB1, H_B2, H_OMB2, H_EMA_M, H_EMA_OM, H_CLIP, H_COUNT };

constexpr int kOptThreads = 256;

// ---- pass 1: sum of squares per chunk (fixed summation order: deterministic) ---------------------------------
__global__ void __launch_bounds__(kOptThreads)
grad_sumsq_kernel(const OptChunk* __restrict__ table, int nchunks, float* __restrict__ chunk_sumsq) {
  pdl_wait();
  __shared__ float red[kOptThreads / 32];
  const int chunk = blockIdx.x;
  const OptChunk c = table[chunk];
  float acc = 0.f;
  if (c.g != nullptr) {
    const float* __restrict__ g = c.g;
    if ((((uintptr_t)g) & 15u) == 0) {
      const int n4 = c.n >> 2;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int i = threadIdx.x; i < n4; i += kOptThreads) {
        const float4 x = ld_stream_f4(reinterpret_cast<const float4*>(g) + i);
        a0 = fmaf(x.x, x.x, a0); a1 = fmaf(x.y, x.y, a1); a2 = fmaf(x.z, x.z, a2); a3 = fmaf(x.w, x.w, a3);
      }
      acc = (a0 + a1) + (a2 + a3);
      for (int j = (n4 << 2) + threadIdx.x; j < c.n; j += kOptThreads) acc = fmaf(g[j], g[j], acc);
    } else {
      for (int j = threadIdx.x; j < c.n; j += kOptThreads) acc = fmaf(g[j], g[j], acc);
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kOptThreads / 32; ++w) s += red[w];
    chunk_sumsq[chunk] = s;
  }
}

// per tensor: norm = sqrt(sum of its chunks), coef = clip/(norm + 1e-6) if that is < 1, else 1 (utils.py:137-140)
__global__ void clip_coef_kernel(const float* __restrict__ chunk_sumsq, const int* __restrict__ first_chunk, int ntensors,
                                 const float* __restrict__ hyper, float* __restrict__ norms, float* __restrict__ coef) {
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ntensors) return;
  float s = 0.f;
  for (int c = first_chunk[t]; c < first_chunk[t + 1]; ++c) s += chunk_sumsq[c];
  const float nrm = sqrtf(s);
  norms[t] = nrm;
  const float clip = hyper[H_CLIP];
  float cf = 1.f;
  if (clip > 0.f) {
    const float q = __fdiv_rn(clip, __fadd_rn(nrm, 1e-6f));
    if (q < 1.f) cf = q;
  }
  coef[t] = cf;
}

// ---- pass 2 ---------------------------------------------------------------------------------------------------
struct OptScalars {
  float omb1, b2, omb2, step_size, bc2_sqrt, eps, ema_m, ema_om;
};

__device__ __forceinline__ void adamw1(float& p, float g, float& m, float& v, const OptScalars& s, float coef, float decay) {
  g = __fmul_rn(g, coef);                                                   // p.grad.mul_(clip_coef) (1.0 when not clipped)
  p = __fmul_rn(p, decay);                                                  // param.mul_(1 - lr*wd)
  m = __fadd_rn(m, __fmul_rn(__fsub_rn(g, m), s.omb1));                     // exp_avg.lerp_(grad, 1 - beta1)
  v = __fadd_rn(__fmul_rn(v, s.b2), __fmul_rn(__fmul_rn(g, g), s.omb2));    // exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2)
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), s.bc2_sqrt), s.eps);
  p = __fsub_rn(p, __fmul_rn(s.step_size, __fdiv_rn(m, denom)));            // param.addcdiv_(exp_avg, denom, -step_size)
}
__device__ __forceinline__ float ema_of(float k, float q, const OptScalars& s) {
  return __fadd_rn(__fmul_rn(k, s.ema_m), __fmul_rn(q, s.ema_om));
}

__global__ void __launch_bounds__(kOptThreads)
adamw_ema_kernel(const OptChunk* __restrict__ table, int nchunks, const float* __restrict__ hyper,
                 const float* __restrict__ coef, const uint8_t* __restrict__ regularized) {
  pdl_wait();
  OptScalars s;
  s.omb1 = hyper[H_OMB1]; s.b2 = hyper[H_B2]; s.omb2 = hyper[H_OMB2];
  s.step_size = hyper[H_STEP_SIZE]; s.bc2_sqrt = hyper[H_BC2_SQRT]; s.eps = hyper[H_EPS];
  s.ema_m = hyper[H_EMA_M]; s.ema_om = hyper[H_EMA_OM];
  const float decay_reg = hyper[H_DECAY];
  for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
    const OptChunk c = table[chunk];
    const float cf = coef[c.tensor];
    const float decay = regularized[c.tensor] ? decay_reg : 1.f;
    const bool has_g = c.g != nullptr, has_k = c.k != nullptr;
    float* __restrict__ p = c.p;
    const float* __restrict__ g = c.g;
    float* __restrict__ m = c.m;
    float* __restrict__ v = c.v;
    float* __restrict__ k = c.k;
    const uintptr_t al = (uintptr_t)p | (has_g ? ((uintptr_t)g | (uintptr_t)m | (uintptr_t)v) : 0) | (has_k ? (uintptr_t)k : 0);
    if ((al & 15u) == 0) {
      const int n4 = c.n >> 2;
      // two 128-bit loads per operand in flight per thread (5 operands: 10 loads)
      for (int i = threadIdx.x; i < n4; i += 2 * kOptThreads) {
        const int i1 = i + kOptThreads;
        const bool two = i1 < n4;
        float4 pv[2], gv[2], mv[2], vv[2], kv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int idx = u == 0 ? i : i1;
          if (u == 0 || two) {
            pv[u] = ld_stream_f4(reinterpret_cast<const float4*>(p) + idx);
            if (has_g) {
              gv[u] = ld_stream_f4(reinterpret_cast<const float4*>(g) + idx);
              mv[u] = ld_stream_f4(reinterpret_cast<const float4*>(m) + idx);
              vv[u] = ld_stream_f4(reinterpret_cast<const float4*>(v) + idx);
            }
            if (has_k) kv[u] = ld_stream_f4(reinterpret_cast<const float4*>(k) + idx);
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int idx = u == 0 ? i : i1;
          if (u == 0 || two) {
            if (has_g) {
              adamw1(pv[u].x, gv[u].x, mv[u].x, vv[u].x, s, cf, decay);
              adamw1(pv[u].y, gv[u].y, mv[u].y, vv[u].y, s, cf, decay);
              adamw1(pv[u].z, gv[u].z, mv[u].z, vv[u].z, s, cf, decay);
              adamw1(pv[u].w, gv[u].w, mv[u].w, vv[u].w, s, cf, decay);
              st_stream_f4(reinterpret_cast<float4*>(p) + idx, pv[u]);
              st_stream_f4(reinterpret_cast<float4*>(m) + idx, mv[u]);
              st_stream_f4(reinterpret_cast<float4*>(v) + idx, vv[u]);
            }
            if (has_k) {
              float4 r;
              r.x = ema_of(kv[u].x, pv[u].x, s); r.y = ema_of(kv[u].y, pv[u].y, s);
              r.z = ema_of(kv[u].z, pv[u].z, s); r.w = ema_of(kv[u].w, pv[u].w, s);
              st_stream_f4(reinterpret_cast<float4*>(k) + idx, r);
            }
          }
        }
      }
      for (int j = (n4 << 2) + threadIdx.x; j < c.n; j += kOptThreads) {
        float pj = p[j];
        if (has_g) { float mj = m[j], vj = v[j]; adamw1(pj, g[j], mj, vj, s, cf, decay); p[j] = pj; m[j] = mj; v[j] = vj; }
        if (has_k) k[j] = ema_of(k[j], pj, s);
      }
    } else {
      for (int j = threadIdx.x; j < c.n; j += kOptThreads) {
        float pj = p[j];
        if (has_g) { float mj = m[j], vj = v[j]; adamw1(pj, g[j], mj, vj, s, cf, decay); p[j] = pj; m[j] = mj; v[j] = vj; }
        if (has_k) k[j] = ema_of(k[j], pj, s);
      }
    }
  }
}

}  // namespace lafs

using namespace lafs;

extern "C" size_t lafs_optim_workspace_bytes(int nchunks, int ntensors) {
  if (nchunks <= 0 || ntensors <= 0) return 0;
  return ((size_t)nchunks * sizeof(float) + 255) / 256 * 256;
}

extern "C" int lafs_adamw_ema_multi(const void* table, int nchunks, const int* first_chunk, const unsigned char* regularized,
                                    int ntensors, const float* hyper, float* grad_norms, float* clip_coef, void* workspace,
                                    size_t workspace_bytes, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(table)) return brc;
  if (nchunks == 0) return LAFS_OK;
  LAFS_REQUIRE(table && first_chunk && regularized && hyper && grad_norms && clip_coef && workspace, LAFS_ERR_ARG,
               "lafs_adamw_ema_multi: null pointer");
  LAFS_REQUIRE(nchunks > 0 && ntensors > 0, LAFS_ERR_ARG, "lafs_adamw_ema_multi: nchunks=%d ntensors=%d", nchunks, ntensors);
  LAFS_REQUIRE(workspace_bytes >= lafs_optim_workspace_bytes(nchunks, ntensors), LAFS_ERR_WORKSPACE,
               "lafs_adamw_ema_multi: workspace %zu too small", workspace_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  float* chunk_sumsq = (float*)workspace;
  launch_pdl((grad_sumsq_kernel), dim3(nchunks), dim3(kOptThreads), (size_t)(0), st, reinterpret_cast<const OptChunk*>(table), nchunks, chunk_sumsq);
  launch_pdl((clip_coef_kernel), dim3((ntensors + 127) / 128), dim3(128), (size_t)(0), st, chunk_sumsq, first_chunk, ntensors, hyper, grad_norms, clip_coef);
  launch_pdl((adamw_ema_kernel), dim3(nchunks), dim3(kOptThreads), (size_t)(0), st, reinterpret_cast<const OptChunk*>(table), nchunks, hyper, clip_coef, regularized);
  return check_launch("lafs_adamw_ema_multi");
}
