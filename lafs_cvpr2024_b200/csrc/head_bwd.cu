// (4) Margin head backward: the two gradient GEMMs on tcgen05 tensor cores.
//   G [B, C_local] bf16 = (softmax - target) * s * grad_out / B      (lafs_head_grad_logits)
//   dE_hat [B, D]       = G   . W_hat      (K = classes, split-K over CTAs, fp32 partials)
//   dW_hat [C_local, D] = G^T . E_hat      (K = batch)
// followed by the F.normalize Jacobians (d x_hat / d x = (I - x_hat x_hat^T)/||x||).
// The reference gets these from autograd through CosFace.forward (ViT_face.py:49-89) and the
// loss; here nothing of size [B, C] is kept in fp32.
//
// One generic kernel: C[M,N] (+)= A . B with B always MN-major in shared memory (its global
// tensor is [K rows, N cols] row-major) and A either K-major ([M rows, K cols] row-major) or
// MN-major ([K rows, M cols] row-major).  MN-major operands use the 128-byte-swizzle canonical
// layout: atoms of 64 (MN, contiguous 128 B) x 8 (K rows), SBO = 1024 B between K groups,
// LBO = bytes between consecutive 64-wide MN blocks; TMA boxes {64 MN, 64 K} land in exactly
// that layout.
#include "umma.cuh"
#include "../../include/lafs_b200.h"

namespace lafs {
using namespace umma;

namespace gb {
constexpr int kBM = 128, kBN = 256, kBK = 64;
constexpr int kStageA = kBM * kBK * 2;      // 16 KB
constexpr int kStageB = kBN * kBK * 2;      // 32 KB
constexpr int kStages = 4;
constexpr int kThreads = 192;
constexpr int kPitch = 36;                              // floats per staged row (144 B: 16-byte aligned, conflict-free)
constexpr int kStageOut = 4 * 32 * kPitch * 4;          // per epilogue warp a [32 x 32] fp32 block: 18,432 B total
constexpr int kSmem = kStages * (kStageA + kStageB) + kStageOut + 1024 + 256;
}  // namespace gb

// MN-major SW128 descriptor: LBO = distance between 64-element MN blocks, SBO = 1024 B
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct GemmParams {
  int M, N, K;            // logical sizes
  int m_tiles, n_tiles, splits, kblocks_per_split, kblocks_total;
  float* out;             // [splits][M][ldo] fp32
  long long ldo, split_stride;
};

template <bool A_MN>
__global__ void __launch_bounds__(gb::kThreads, 1)
gemm_bwd_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const GemmParams p) {
  using namespace gb;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_a = smem;
  uint8_t* s_b = smem + kStages * kStageA;
  float* s_out = reinterpret_cast<float*>(s_b + kStages * kStageB);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_b + kStages * kStageB + kStageOut);
  uint64_t* full = bars;                 // kStages
  uint64_t* empty = bars + kStages;      // kStages
  uint64_t* acc_full = bars + 2 * kStages;      // 2
  uint64_t* acc_empty = acc_full + 2;           // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_b);
    for (int i = 0; i < kStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * kBN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int split = tile % p.splits;
        const int mn = tile / p.splits;
        const int nt = mn % p.n_tiles, mt = mn / p.n_tiles;
        const int kb0 = split * p.kblocks_per_split;
        const int kb1 = min(kb0 + p.kblocks_per_split, p.kblocks_total);
        for (int kb = kb0; kb < kb1; ++kb, ++cnt) {
          const int st = cnt % kStages;
          mbar_wait(empty + st, ((cnt / kStages) & 1) ^ 1);
          mbar_arrive_expect_tx(full + st, kStageA + kStageB);
          uint8_t* da = s_a + st * kStageA;
          uint8_t* db = s_b + st * kStageB;
          if (A_MN) {   // A global [K rows, M cols]: two boxes {64 m, 64 k}
            tma_load_2d(da, &tmap_a, full + st, mt * kBM, kb * kBK);
            tma_load_2d(da + 8192, &tmap_a, full + st, mt * kBM + 64, kb * kBK);
          } else {      // A global [M rows, K cols]: one box {64 k, 128 m}
            tma_load_2d(da, &tmap_a, full + st, kb * kBK, mt * kBM);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)   // B global [K rows, N cols]: four boxes {64 n, 64 k}
            tma_load_2d(db + q * 8192, &tmap_b, full + st, nt * kBN + q * 64, kb * kBK);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBM, kBN, A_MN ? 1 : 0, 1);
      uint32_t cnt = 0, acnt = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acnt) {
        const int split = tile % p.splits;
        const int kb0 = split * p.kblocks_per_split;
        const int kb1 = min(kb0 + p.kblocks_per_split, p.kblocks_total);
        const int buf = acnt & 1;
        mbar_wait(acc_empty + buf, ((acnt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * kBN);
        for (int kb = kb0; kb < kb1; ++kb, ++cnt) {
          const int st = cnt % kStages;
          mbar_wait(full + st, (cnt / kStages) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(s_a + st * kStageA);
          const uint32_t b_addr = smem_u32(s_b + st * kStageB);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            // 16 K per UMMA: K-major advances 32 B inside the swizzle atom; MN-major advances
            // two 8-row K groups (2 x 1024 B)
            const uint64_t da = A_MN ? make_desc_mn_sw128(a_addr + kk * 2048, 8192)
                                     : desc_advance_k(make_desc_k_sw128(a_addr), kk * 16);
            const uint64_t db = make_desc_mn_sw128(b_addr + kk * 2048, 8192);
            mma_f16_ss(d_tmem, da, db, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
          }
          mma_commit(empty + st);
        }
        mma_commit(acc_full + buf);
      }
    }
  } else {
    const int quarter = warp & 3;
    uint32_t acnt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acnt) {
      const int split = tile % p.splits;
      const int mn = tile / p.splits;
      const int nt = mn % p.n_tiles, mt = mn / p.n_tiles;
      const int buf = acnt & 1;
      mbar_wait(acc_full + buf, (acnt >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(buf * kBN) + ((uint32_t)(quarter * 32) << 16);
      // lane = output row in TMEM, but rows are ldo floats apart in memory: transpose each
      // [32 rows x 32 cols] block through shared memory so that a warp-wide store writes four full
      // 128-byte row segments instead of 32 scattered 16-byte pieces
      float* stage = s_out + quarter * (32 * kPitch);
      const int row_base = mt * kBM + quarter * 32;
      float* dst0 = p.out + (size_t)split * p.split_stride + (size_t)nt * kBN;
#pragma unroll 1
      for (int piece = 0; piece < kBN / 32; ++piece) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + (uint32_t)(piece * 32), v);
        tmem_ld_wait();
        const int c0 = nt * kBN + piece * 32;
        if (c0 >= p.N) break;                       // warp-uniform
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(stage + lane * kPitch + j) =
              make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                          __uint_as_float(v[j + 3]));
        __syncwarp();
#pragma unroll
        for (int r8 = 0; r8 < 8; ++r8) {
          const int rr = r8 * 4 + (lane >> 3), part = lane & 7;       // 8 lanes cover one 128-byte row segment
          const float4 val = *reinterpret_cast<const float4*>(stage + rr * kPitch + part * 4);
          const int grow = row_base + rr, gcol = c0 + part * 4;
          if (grow < p.M) {
            float* d = dst0 + (size_t)grow * p.ldo + piece * 32 + part * 4;
            if (gcol + 4 <= p.N) *reinterpret_cast<float4*>(d) = val;
            else {
              if (gcol < p.N) d[0] = val.x;
              if (gcol + 1 < p.N) d[1] = val.y;
              if (gcol + 2 < p.N) d[2] = val.z;
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kBN);
  }
}

// out[r, :] = (g[r, :] - x_hat[r, :] * <x_hat[r, :], g[r, :]>) * inv_norm[r]   (F.normalize backward)
// one warp per row, 128-bit accesses, all loads of a row issued before the reduction.
// D % 64 == 0, D <= 768: a lane holds up to three (float4 x 2) groups.  out may alias g.
__global__ void __launch_bounds__(256)
normalize_bwd_kernel(const float* __restrict__ g, const __nv_bfloat16* __restrict__ x_hat,
                     const float* __restrict__ inv_norm, int R, int D, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= R) return;
  const float* grow = g + (size_t)r * D;
  const __nv_bfloat16* xrow = x_hat + (size_t)r * D;
  float4 gv[3][2];
  uint4 xv[3];
  // group q covers columns q*256 + lane*8 .. +7
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int c = q * 256 + lane * 8;
    if (c < D) {
      gv[q][0] = *reinterpret_cast<const float4*>(grow + c);
      gv[q][1] = *reinterpret_cast<const float4*>(grow + c + 4);
      xv[q] = *reinterpret_cast<const uint4*>(xrow + c);
    }
  }
  float dot = 0.f;
  float xf[3][8];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int c = q * 256 + lane * 8;
    if (c < D) {
      xf[q][0] = Half2Ops<__nv_bfloat16>::lo(xv[q].x); xf[q][1] = Half2Ops<__nv_bfloat16>::hi(xv[q].x);
      xf[q][2] = Half2Ops<__nv_bfloat16>::lo(xv[q].y); xf[q][3] = Half2Ops<__nv_bfloat16>::hi(xv[q].y);
      xf[q][4] = Half2Ops<__nv_bfloat16>::lo(xv[q].z); xf[q][5] = Half2Ops<__nv_bfloat16>::hi(xv[q].z);
      xf[q][6] = Half2Ops<__nv_bfloat16>::lo(xv[q].w); xf[q][7] = Half2Ops<__nv_bfloat16>::hi(xv[q].w);
      dot += gv[q][0].x * xf[q][0] + gv[q][0].y * xf[q][1] + gv[q][0].z * xf[q][2] + gv[q][0].w * xf[q][3] +
             gv[q][1].x * xf[q][4] + gv[q][1].y * xf[q][5] + gv[q][1].z * xf[q][6] + gv[q][1].w * xf[q][7];
    }
  }
  dot = warp_sum(dot);
  const float inv = inv_norm[r];
  float* orow = out + (size_t)r * D;
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int c = q * 256 + lane * 8;
    if (c < D) {
      float4 o0, o1;
      o0.x = (gv[q][0].x - xf[q][0] * dot) * inv; o0.y = (gv[q][0].y - xf[q][1] * dot) * inv;
      o0.z = (gv[q][0].z - xf[q][2] * dot) * inv; o0.w = (gv[q][0].w - xf[q][3] * dot) * inv;
      o1.x = (gv[q][1].x - xf[q][4] * dot) * inv; o1.y = (gv[q][1].y - xf[q][5] * dot) * inv;
      o1.z = (gv[q][1].z - xf[q][6] * dot) * inv; o1.w = (gv[q][1].w - xf[q][7] * dot) * inv;
      *reinterpret_cast<float4*>(orow + c) = o0;
      *reinterpret_cast<float4*>(orow + c + 4) = o1;
    }
  }
}

__global__ void split_reduce_kernel(const float* __restrict__ part, int splits, long long stride, long long n,
                                    float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float acc = 0.f;
  for (int s = 0; s < splits; ++s) acc += part[(size_t)s * stride + i];
  out[i] = acc;
}

static int encode_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_bytes,
                          uint32_t box_cols, uint32_t box_rows) {
  TmaEncoder::EncodeTiled enc = TmaEncoder::get();
  LAFS_REQUIRE(enc != nullptr, LAFS_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LAFS_REQUIRE(r == CUDA_SUCCESS, LAFS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return LAFS_OK;
}

template <bool A_MN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  auto kern = gemm_bwd_kernel<A_MN>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gb::kSmem);
  LAFS_REQUIRE(e == cudaSuccess, LAFS_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int total = p.m_tiles * p.n_tiles * p.splits;
  const int grid = total < kNumSMs ? total : kNumSMs;
  kern<<<grid, gb::kThreads, gb::kSmem, st>>>(ta, tb, p);
  return check_launch("gemm_bwd_kernel");
}

static int de_splits(int B, int C_local, int D) {
  const int tiles = ((B + 127) / 128) * ((D + 255) / 256);
  const int kblocks = (C_local + 63) / 64;
  int s = kNumSMs / tiles;
  if (s < 1) s = 1;
  if (s > kblocks) s = kblocks;
  return s;
}

}  // namespace lafs

using namespace lafs;

extern "C" size_t lafs_head_bwd_workspace_bytes(int B, int C_local, int D) {
  if (B <= 0 || C_local <= 0 || D <= 0) return 0;
  return (size_t)de_splits(B, C_local, D) * B * D * sizeof(float);
}

/* dE_hat partial sums: G [B,C_local] . W_hat [C_local,D]  (K-major A, split-K), reduced over the
 * splits into grad_e_hat [B,D] fp32 (no Jacobian: sharded heads all-reduce it first). */
extern "C" int lafs_head_bwd_embed(const void* grad_bf16, long long ldg, const void* w_hat, int B, int C_local, int D,
                                   float* grad_e_hat, void* workspace, size_t workspace_bytes, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(grad_bf16)) return brc;
  LAFS_REQUIRE(grad_bf16 && w_hat && workspace && grad_e_hat, LAFS_ERR_ARG, "lafs_head_bwd_embed: null pointer");
  LAFS_REQUIRE(B > 0 && C_local > 0 && D > 0 && D % 64 == 0 && D <= 768, LAFS_ERR_ARG, "lafs_head_bwd_embed: B=%d C=%d D=%d", B, C_local, D);
  LAFS_REQUIRE(ldg % 8 == 0 && ldg >= C_local, LAFS_ERR_ARG, "lafs_head_bwd_embed: ldg=%lld must be a multiple of 8 and >= C_local", ldg);
  const int splits = de_splits(B, C_local, D);
  const size_t need = (size_t)splits * B * D * sizeof(float);
  LAFS_REQUIRE(workspace_bytes >= need, LAFS_ERR_WORKSPACE, "lafs_head_bwd_embed: workspace %zu < %zu", workspace_bytes, need);
  CUtensorMap ta, tb;
  int rc = encode_bf16_2d(&ta, grad_bf16, (uint64_t)B, (uint64_t)C_local, (uint64_t)ldg * 2, 64, 128);  // K-major A
  if (rc) return rc;
  rc = encode_bf16_2d(&tb, w_hat, (uint64_t)C_local, (uint64_t)D, (uint64_t)D * 2, 64, 64);            // MN-major B
  if (rc) return rc;
  GemmParams p{};
  p.M = B; p.N = D; p.K = C_local;
  p.m_tiles = (B + 127) / 128; p.n_tiles = (D + 255) / 256; p.splits = splits;
  p.kblocks_total = (C_local + 63) / 64;
  p.kblocks_per_split = (p.kblocks_total + splits - 1) / splits;
  p.splits = (p.kblocks_total + p.kblocks_per_split - 1) / p.kblocks_per_split;   // no empty K ranges
  p.out = (float*)workspace; p.ldo = D; p.split_stride = (long long)B * D;
  cudaStream_t st = (cudaStream_t)stream;
  rc = launch_gemm<false>(ta, tb, p, st);
  if (rc) return rc;
  const long long n = (long long)B * D;
  split_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const float*)workspace, p.splits, n, n, grad_e_hat);
  return check_launch("lafs_head_bwd_embed");
}

/* dW_hat [C_local,D] = G^T . E_hat (both operands MN-major), then the normalisation Jacobian of the
 * weight rows in place:  grad_w = (dW_hat - w_hat <w_hat, dW_hat>) * inv_norm_w. */
extern "C" int lafs_head_bwd_weight(const void* grad_bf16, long long ldg, const void* e_hat, const void* w_hat,
                                    const float* inv_norm_w, int B, int C_local, int D, float* grad_w,
                                    lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(grad_bf16)) return brc;
  LAFS_REQUIRE(grad_bf16 && e_hat && w_hat && inv_norm_w && grad_w, LAFS_ERR_ARG, "lafs_head_bwd_weight: null pointer");
  LAFS_REQUIRE(B > 0 && C_local > 0 && D > 0 && D % 64 == 0 && D <= 768, LAFS_ERR_ARG, "lafs_head_bwd_weight: B=%d C=%d D=%d", B, C_local, D);
  LAFS_REQUIRE(ldg % 8 == 0 && ldg >= C_local, LAFS_ERR_ARG, "lafs_head_bwd_weight: ldg=%lld must be a multiple of 8 and >= C_local", ldg);
  CUtensorMap ta, tb;
  int rc = encode_bf16_2d(&ta, grad_bf16, (uint64_t)B, (uint64_t)C_local, (uint64_t)ldg * 2, 64, 64);  // MN-major A: [K=B, M=C]
  if (rc) return rc;
  rc = encode_bf16_2d(&tb, e_hat, (uint64_t)B, (uint64_t)D, (uint64_t)D * 2, 64, 64);                  // MN-major B: [K=B, N=D]
  if (rc) return rc;
  GemmParams p{};
  p.M = C_local; p.N = D; p.K = B;
  p.m_tiles = (C_local + 127) / 128; p.n_tiles = (D + 255) / 256; p.splits = 1;
  p.kblocks_total = (B + 63) / 64; p.kblocks_per_split = p.kblocks_total;
  p.out = grad_w; p.ldo = D; p.split_stride = 0;
  cudaStream_t st = (cudaStream_t)stream;
  rc = launch_gemm<true>(ta, tb, p, st);
  if (rc) return rc;
  normalize_bwd_kernel<<<(C_local + 7) / 8, 256, 0, st>>>(grad_w, (const __nv_bfloat16*)w_hat, inv_norm_w, C_local, D, grad_w);
  return check_launch("lafs_head_bwd_weight");
}

/* out[r,:] = (g[r,:] - x_hat[r,:] <x_hat[r,:], g[r,:]>) * inv_norm[r]   (F.normalize backward); out may alias g */
extern "C" int lafs_normalize_bwd(const float* g, const void* x_hat_bf16, const float* inv_norm, int R, int D, float* out,
                                  lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(g)) return brc;
  LAFS_REQUIRE(g && x_hat_bf16 && inv_norm && out && R >= 0 && D > 0 && D <= 768, LAFS_ERR_ARG, "lafs_normalize_bwd: bad argument");
  if (R == 0) return LAFS_OK;
  LAFS_REQUIRE(D % 8 == 0, LAFS_ERR_ARG, "lafs_normalize_bwd: D=%d must be a multiple of 8", D);
  normalize_bwd_kernel<<<(R + 7) / 8, 256, 0, (cudaStream_t)stream>>>(g, (const __nv_bfloat16*)x_hat_bf16, inv_norm, R, D, out);
  return check_launch("lafs_normalize_bwd");
}
