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
#include <stdlib.h>
#include "umma.cuh"
#include "../../include/lafs_b200.h"

namespace lafs {
using namespace umma;

namespace gb {
constexpr int kBM = 128, kBN = 256, kBK = 64;
constexpr int kStageA = kBM * kBK * 2;      // 16 KB
constexpr int kStageB = kBN * kBK * 2;      // 32 KB
constexpr int kStages = 4;
constexpr int kPrefetch = 8;                             // k blocks of L2 prefetch distance
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

// CL > 1: the CL CTAs of a cluster work on CL consecutive M tiles of the same (N tile, K split) in
// lock step and share the B operand: each fetches 1/CL of every B k block (one {64 n, 64 k} box of
// four) and multicasts it to the cluster, so the L2 -> SM fill of B drops by CL (dE: every M tile
// contracts against the same W_hat range, and that fill is what bounds the single-CTA kernel).
template <bool A_MN, int CL>
__global__ void __launch_bounds__(gb::kThreads, 1)
gemm_bwd_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const GemmParams p) {
  pdl_wait();
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
    for (int i = 0; i < kStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, CL); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * kBN);
  tc_fence_before();
  if (CL > 1) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item = (M group of CL tiles, N tile, K split); CTA `rank` of the cluster takes M tile
  // group * CL + rank (tiles past the end load zeros and store nothing)
  const int rank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int first = blockIdx.x / CL, stride = gridDim.x / CL;
  const int m_groups = (p.m_tiles + CL - 1) / CL;
  const int total_tiles = m_groups * p.n_tiles * p.splits;
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);

  if (warp == 0) {
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int tile = first; tile < total_tiles; tile += stride) {
        const int split = tile % p.splits;
        const int mn = tile / p.splits;
        const int nt = mn % p.n_tiles, mt = (mn / p.n_tiles) * CL + rank;
        const int kb0 = split * p.kblocks_per_split;
        const int kb1 = min(kb0 + p.kblocks_per_split, p.kblocks_total);
        for (int kb = kb0; kb < kb1; ++kb, ++cnt) {
          const int st = cnt % kStages;
          // pull the operands of k block kb + kPrefetch into L2 (once per operand tile: A by the N tile 0
          // CTAs, B by the M tile 0 CTAs): the 4-stage ring alone does not cover HBM latency
          if (kb + kPrefetch < kb1) {
            const int kp = (kb + kPrefetch) * kBK;
            if (nt == 0) {
              if (A_MN) { tma_prefetch_l2_2d(&tmap_a, mt * kBM, kp); tma_prefetch_l2_2d(&tmap_a, mt * kBM + 64, kp); }
              else tma_prefetch_l2_2d(&tmap_a, kp, mt * kBM);
            }
            if (mt == 0) {
#pragma unroll
              for (int q = 0; q < 4; ++q) tma_prefetch_l2_2d(&tmap_b, nt * kBN + q * 64, kp);
            }
          }
          mbar_wait(empty + st, ((cnt / kStages) & 1) ^ 1);   // all CL consumers of this slot are done
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
          for (int q = 0; q < 4; ++q) {   // B global [K rows, N cols]: four boxes {64 n, 64 k}
            if (CL == 1) tma_load_2d(db + q * 8192, &tmap_b, full + st, nt * kBN + q * 64, kb * kBK);
            else if (q % CL == rank) tma_load_2d_mc(db + q * 8192, &tmap_b, full + st, nt * kBN + q * 64, kb * kBK, kMask);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t cnt = 0, acnt = 0;
      for (int tile = first; tile < total_tiles; tile += stride, ++acnt) {
        const int split = tile % p.splits;
        // UMMA N = the valid columns of this N tile rounded up to 16 (a [*, 208] problem issues N = 208, not 256)
        const int n_left = p.N - ((tile / p.splits) % p.n_tiles) * kBN;
        const uint32_t idesc = make_idesc_bf16(kBM, n_left >= kBN ? kBN : ((n_left + 15) & ~15), A_MN ? 1 : 0, 1);
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
          if (CL > 1) mma_commit_mc(empty + st, kMask);
          else mma_commit(empty + st);
        }
        mma_commit(acc_full + buf);
      }
    }
  } else {
    const int quarter = warp & 3;
    uint32_t acnt = 0;
    for (int tile = first; tile < total_tiles; tile += stride, ++acnt) {
      const int split = tile % p.splits;
      const int mn = tile / p.splits;
      const int nt = mn % p.n_tiles, mt = (mn / p.n_tiles) * CL + rank;
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
      if (lane == 0) mbar_arrive_relaxed(acc_empty + buf);
    }
  }
  tc_fence_before();
  if (CL > 1) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * kBN);
  }
}

// (A single-kernel dW with the F.normalize Jacobian applied in place from L2 by the same CTA -- full 128 x 512 rows
//  in TMEM, optional E_hat multicast over clusters -- was built and measured in rounds 1-2 and removed: 157.8 us
//  against 138.2 us for GEMM + reverse-order Jacobian pass at B=512, C=93431, D=512; its in-place phase is a chain of
//  L2 round trips.  The form that won is dw_diag_kernel below.)

// ---- dW with the Jacobian's rank-one term on the tensor core (the default dW path) ---------------------------------
//   grad_w[c,:] = inv_norm_w[c] * ( dW_hat[c,:] - t[c] * w_hat[c,:] ),   t[c] = <w_hat[c,:], dW_hat[c,:]> = sum_b G[b,c]*cos[b,c]
// t comes from the gradient kernel (HEAD_GRAD_T partials).  Per 128-class tile the correction is the product
// (-diag(t)) [128 x 128] . W_hat_tile [128 x D]: two extra k blocks of the same GEMM, whose A operand (a
// diagonal matrix, bf16) is written into the stage by the producer warp in the MN-major / 128-byte-swizzle
// layout TMA would produce, and whose B operand is the W_hat tile itself, fetched by TMA like E_hat.  The
// epilogue only scales each row by inv_norm_w: no second pass over dW, no lane-per-row loads of w_hat.
// (fp32 emulation: 6e-4 max-norm relative error on dW from rounding t to bf16.)
struct DiagParams {
  const float* t;           // [M] = sum of the gradient kernel's partial rows (t_reduce_kernel)
  const float* inv_norm;    // [M]
};

// t[c] = sum_q tpart[q][c], in place into row 0 (a thread owns a column: no hazard)
__global__ void t_reduce_kernel(float* __restrict__ tpart, int tparts, long long ldt, int C) {
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float acc = 0.f;
  for (int q = 0; q < tparts; ++q) acc += tpart[(size_t)q * ldt + c];
  tpart[c] = acc;
}

__global__ void __launch_bounds__(gb::kThreads, 1)
dw_diag_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_w, const GemmParams p, const DiagParams dp) {
  pdl_wait();
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
    prefetch_tensormap(&tmap_w);
    for (int i = 0; i < kStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 2 * kBN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int total_tiles = p.m_tiles * p.n_tiles;
  const int nkb = p.kblocks_total;            // batch k blocks; + 2 diagonal blocks per tile

  if (warp == 0) {
    uint32_t cnt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
      // this lane's four diagonal entries (classes lane, lane+32 of either 64-class half), fetched before the
      // batch blocks so that their latency is hidden behind them
      float tv[2][2];
#pragma unroll
      for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int cls = mt * kBM + e * 64 + lane + 32 * h;
          tv[e][h] = cls < p.M ? __ldg(dp.t + cls) : 0.f;
        }
      if (lane == 0) {
        for (int kb = 0; kb < nkb; ++kb) {
          const uint32_t c = cnt + (uint32_t)kb;
          const int st = c % kStages;
          mbar_wait(empty + st, ((c / kStages) & 1) ^ 1);
          mbar_arrive_expect_tx(full + st, kStageA + kStageB);
          uint8_t* da = s_a + st * kStageA;
          uint8_t* db = s_b + st * kStageB;
          tma_load_2d(da, &tmap_a, full + st, mt * kBM, kb * kBK);             // G^T: {64 classes, 64 batch rows}
          tma_load_2d(da + 8192, &tmap_a, full + st, mt * kBM + 64, kb * kBK);
#pragma unroll
          for (int q = 0; q < 4; ++q) tma_load_2d(db + q * 8192, &tmap_b, full + st, nt * kBN + q * 64, kb * kBK);
        }
      }
      cnt += (uint32_t)nkb;
      __syncwarp();
      // two diagonal blocks: K index = class within the tile, e*64 .. e*64+63
      for (int e = 0; e < 2; ++e, ++cnt) {
        const int st = cnt % kStages;
        mbar_wait(empty + st, ((cnt / kStages) & 1) ^ 1);      // every lane waits: all of them write the stage
        uint8_t* da = s_a + st * kStageA;
        uint8_t* db = s_b + st * kStageB;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int i = lane; i < kStageA / 16; i += 32) reinterpret_cast<uint4*>(da)[i] = z;
        __syncwarp();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int kk = lane + 32 * h;                          // row k = kk of the block, column m' = kk of box e
          const uint32_t off = (uint32_t)(e * 8192 + kk * 128 + ((((kk >> 3) ^ (kk & 7)) & 7) << 4) + (kk & 7) * 2);
          *reinterpret_cast<__nv_bfloat16*>(da + off) = __float2bfloat16_rn(-(e == 0 ? tv[0][h] : tv[1][h]));
        }
        fence_proxy_async_smem();        // generic-proxy writes -> visible to the UMMA operand reads
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_expect_tx(full + st, kStageB);
#pragma unroll
          for (int q = 0; q < 4; ++q)    // W_hat tile rows as the B operand: {64 d, 64 classes}
            tma_load_2d(db + q * 8192, &tmap_w, full + st, nt * kBN + q * 64, mt * kBM + e * 64);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBM, kBN, 1, 1);
      uint32_t cnt = 0, acnt = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++acnt) {
        const int buf = acnt & 1;
        mbar_wait(acc_empty + buf, ((acnt >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * kBN);
        for (int kb = 0; kb < nkb + 2; ++kb, ++cnt) {
          const int st = cnt % kStages;
          mbar_wait(full + st, (cnt / kStages) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(s_a + st * kStageA);
          const uint32_t b_addr = smem_u32(s_b + st * kStageB);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t da = make_desc_mn_sw128(a_addr + kk * 2048, 8192);
            const uint64_t db = make_desc_mn_sw128(b_addr + kk * 2048, 8192);
            mma_f16_ss(d_tmem, da, db, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
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
      const int nt = tile % p.n_tiles, mt = tile / p.n_tiles;
      const int buf = acnt & 1;
      const int row_base = mt * kBM + quarter * 32;
      const float inv = row_base + lane < p.M ? __ldg(dp.inv_norm + row_base + lane) : 0.f;
      mbar_wait(acc_full + buf, (acnt >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(buf * kBN) + ((uint32_t)(quarter * 32) << 16);
      float* stage = s_out + quarter * (32 * kPitch);
      float* dst0 = p.out + (size_t)nt * kBN;
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
              make_float4(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv, __uint_as_float(v[j + 2]) * inv,
                          __uint_as_float(v[j + 3]) * inv);
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
      if (lane == 0) mbar_arrive_relaxed(acc_empty + buf);
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
                     const float* __restrict__ inv_norm, int R, int D, float* __restrict__ out, int reverse) {
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int r = blockIdx.x * 8 + warp;
  if (r >= R) return;
  // reverse: start with the rows the producing GEMM wrote LAST -- they are still in L2 (126 MB), so
  // the first ~100 MB of this pass do not touch HBM
  if (reverse) r = R - 1 - r;
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
  pdl_wait();
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

// co-resident clusters of CL CTAs of gemm_bwd_kernel (0 if the query fails)
template <bool A_MN, int CL>
static int max_clusters_query();
template <bool A_MN, int CL>
static int max_clusters() {
  if (CL == 1) return kNumSMs;
  // a property of (kernel, shared-memory size, device model): queried once per process (one process
  // drives one GPU model; a benign race at worst repeats the query)
  static int cached = -1;
  if (cached >= 0) return cached;
  cached = max_clusters_query<A_MN, CL>();
  return cached;
}
template <bool A_MN, int CL>
static int max_clusters_query() {
  auto kern = gemm_bwd_kernel<A_MN, CL>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gb::kSmem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL);
  cfg.blockDim = dim3(gb::kThreads);
  cfg.dynamicSmemBytes = gb::kSmem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n * CL > kNumSMs ? kNumSMs / CL : n;
}

template <bool A_MN, int CL>
static int launch_gemm_cl(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int nclusters, cudaStream_t st) {
  auto kern = gemm_bwd_kernel<A_MN, CL>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gb::kSmem);
  LAFS_REQUIRE(e == cudaSuccess, LAFS_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(gb::kThreads);
  cfg.dynamicSmemBytes = gb::kSmem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see common.cuh: every kernel starts with pdl_wait()
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  const int total = ((p.m_tiles + CL - 1) / CL) * p.n_tiles * p.splits;
  cfg.gridDim = dim3((total < nclusters ? total : nclusters) * CL);
  e = cudaLaunchKernelEx(&cfg, kern, ta, tb, p);
  LAFS_REQUIRE(e == cudaSuccess, LAFS_ERR_CUDA, "gemm_bwd_kernel launch: %s", cudaGetErrorString(e));
  return check_launch("gemm_bwd_kernel");
}

template <bool A_MN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  return launch_gemm_cl<A_MN, 1>(ta, tb, p, kNumSMs, st);
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
  p.m_tiles = (B + 127) / 128; p.n_tiles = (D + 255) / 256;
  p.kblocks_total = (C_local + 63) / 64;
  p.out = (float*)workspace; p.ldo = D; p.split_stride = (long long)B * D;
  // W_hat multicast width (LAFS_DE_CLUSTER: 1, 2 or 4).  Measured on B200 (round 2): pairs are never slower than
  // quads -- B=1024, C=205990, D=512: 206 us (2) / 222 us (4) / 228 us (1); B=512, C=93431: 64.5 us for all three;
  // the fused DINO head's two calls (1536 / 512 rows, K=65536, D=256): 0.4157 ms (2) / 0.421 ms (4) per step
  int want = 2;
  if (const char* e = getenv("LAFS_DE_CLUSTER")) want = atoi(e);
  int cl = 1, nclusters = kNumSMs;
  if (want >= 4 && p.m_tiles % 4 == 0) {
    const int n4 = max_clusters<false, 4>();
    if (n4 * 4 >= (kNumSMs * 3) / 4) { cl = 4; nclusters = n4; }
  }
  if (cl == 1 && want >= 2 && p.m_tiles % 2 == 0) {
    const int n2 = max_clusters<false, 2>();
    if (n2 * 2 >= (kNumSMs * 3) / 4) { cl = 2; nclusters = n2; }
  }
  int s_max = nclusters / ((p.m_tiles / cl) * p.n_tiles);   // one wave
  if (s_max < 1) s_max = 1;
  if (s_max > splits) s_max = splits;                       // the workspace was sized for `splits`
  if (s_max > p.kblocks_total) s_max = p.kblocks_total;
  p.kblocks_per_split = (p.kblocks_total + s_max - 1) / s_max;
  p.splits = (p.kblocks_total + p.kblocks_per_split - 1) / p.kblocks_per_split;   // no empty K ranges
  cudaStream_t st = (cudaStream_t)stream;
  if (cl == 4) rc = launch_gemm_cl<false, 4>(ta, tb, p, nclusters, st);
  else if (cl == 2) rc = launch_gemm_cl<false, 2>(ta, tb, p, nclusters, st);
  else rc = launch_gemm_cl<false, 1>(ta, tb, p, kNumSMs, st);
  if (rc) return rc;
  const long long n = (long long)B * D;
  launch_pdl((split_reduce_kernel), dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)(0), st, (const float*)workspace, p.splits, n, n, grad_e_hat);
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
  cudaStream_t st = (cudaStream_t)stream;
  GemmParams p{};
  p.M = C_local; p.N = D; p.K = B;
  p.m_tiles = (C_local + 127) / 128; p.n_tiles = (D + 255) / 256; p.splits = 1;
  p.kblocks_total = (B + 63) / 64; p.kblocks_per_split = p.kblocks_total;
  p.out = grad_w; p.ldo = D; p.split_stride = 0;
  rc = launch_gemm<true>(ta, tb, p, st);
  if (rc) return rc;
  launch_pdl((normalize_bwd_kernel), dim3((C_local + 7) / 8), dim3(256), (size_t)(0), st, grad_w, (const __nv_bfloat16*)w_hat, inv_norm_w, C_local, D, grad_w, 1);
  return check_launch("lafs_head_bwd_weight");
}

/* out[r,:] = (g[r,:] - x_hat[r,:] <x_hat[r,:], g[r,:]>) * inv_norm[r]   (F.normalize backward); out may alias g */
extern "C" int lafs_normalize_bwd(const float* g, const void* x_hat_bf16, const float* inv_norm, int R, int D, float* out,
                                  lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(g)) return brc;
  LAFS_REQUIRE(g && x_hat_bf16 && inv_norm && out && R >= 0 && D > 0 && D <= 768, LAFS_ERR_ARG, "lafs_normalize_bwd: bad argument");
  if (R == 0) return LAFS_OK;
  LAFS_REQUIRE(D % 8 == 0, LAFS_ERR_ARG, "lafs_normalize_bwd: D=%d must be a multiple of 8", D);
  launch_pdl((normalize_bwd_kernel), dim3((R + 7) / 8), dim3(256), (size_t)(0), (cudaStream_t)stream, g, (const __nv_bfloat16*)x_hat_bf16, inv_norm, R, D, out, 0);
  return check_launch("lafs_normalize_bwd");
}

/* ------------------------------------------------------------------------------------------------
 * Backward of patch_to_embedding = nn.Linear(192, dim) (ViT_face.py:760-761, lafs_train.py:544) on the
 * same generic tcgen05 GEMM: the training path of the fused gather -> embed kernel.
 *   grad_w [dim,192]  = grad_emb^T [dim, M] . tokens [M, 192]    (both operands MN-major, split-K over M)
 *   grad_tok [M,192]  = grad_emb [M, dim] . W [dim, 192]         (K-major A, MN-major B)
 * M = faces * landmarks token rows; operands bf16, accumulation and outputs fp32. */
static int embed_bwd_splits(int M, int dim) {
  const int tiles = (dim + 127) / 128;
  const int kblocks = (M + 63) / 64;
  int s = kNumSMs / tiles;
  if (s < 1) s = 1;
  if (s > kblocks) s = kblocks;
  return s;
}

constexpr int kTokLd = 208;   // row pitch of the saved-token tensor (patch_embed.cu): 192 features, a ones column, 15 zeros

extern "C" size_t lafs_embed_bwd_workspace_bytes(int M, int dim) {
  if (M <= 0 || dim <= 0) return 0;
  return (size_t)embed_bwd_splits(M, dim) * dim * kTokLd * sizeof(float);
}

// sum of the split-K partials [splits][dim][208] -> grad_w [dim][192] in the reference's feature order
// ((i*8+j)*3 + c  <-  kernel order c*64 + j*8 + i) and grad_b [dim] (the ones column, k = 192)
__global__ void embed_dw_unpermute_kernel(const float* __restrict__ part, int splits, int dim, float* __restrict__ grad_w,
                                          float* __restrict__ grad_b, int accumulate) {
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= dim * 193) return;
  const int d = idx / 193, k = idx - d * 193;
  float acc = 0.f;
  for (int s = 0; s < splits; ++s) acc += part[((size_t)s * dim + d) * kTokLd + k];
  if (k == 192) {
    if (grad_b != nullptr) grad_b[d] = accumulate ? grad_b[d] + acc : acc;
    return;
  }
  const int c = k >> 6, j = (k >> 3) & 7, i = k & 7;
  float* o = grad_w + (size_t)d * 192 + (i * 8 + j) * 3 + c;
  *o = accumulate ? *o + acc : acc;
}

extern "C" int lafs_embed_bwd_weight_perm(const void* grad_emb_bf16, const void* tokens_perm_bf16, int M, int dim, float* grad_w,
                                          float* grad_b, int accumulate, void* workspace, size_t workspace_bytes,
                                          lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(grad_emb_bf16)) return brc;
  LAFS_REQUIRE(grad_emb_bf16 && tokens_perm_bf16 && grad_w && workspace, LAFS_ERR_ARG, "lafs_embed_bwd_weight_perm: null pointer");
  LAFS_REQUIRE(M > 0 && dim > 0 && dim % 8 == 0, LAFS_ERR_ARG, "lafs_embed_bwd_weight_perm: M=%d dim=%d (dim must be a multiple of 8)", M, dim);
  LAFS_REQUIRE((((uintptr_t)grad_emb_bf16 | (uintptr_t)tokens_perm_bf16 | (uintptr_t)workspace) & 15u) == 0, LAFS_ERR_ARG,
               "lafs_embed_bwd_weight_perm: pointers must be 16-byte aligned");
  const int splits = embed_bwd_splits(M, dim);
  const size_t need = (size_t)splits * dim * kTokLd * sizeof(float);
  LAFS_REQUIRE(workspace_bytes >= need, LAFS_ERR_WORKSPACE, "lafs_embed_bwd_weight_perm: workspace %zu < %zu", workspace_bytes, need);
  CUtensorMap ta, tb;
  int rc = encode_bf16_2d(&ta, grad_emb_bf16, (uint64_t)M, (uint64_t)dim, (uint64_t)dim * 2, 64, 64);     // MN-major A: [K=M, dim]
  if (rc) return rc;
  rc = encode_bf16_2d(&tb, tokens_perm_bf16, (uint64_t)M, kTokLd, kTokLd * 2, 64, 64);                    // MN-major B: [K=M, 208]
  if (rc) return rc;
  GemmParams p{};
  p.M = dim; p.N = kTokLd; p.K = M;
  p.m_tiles = (dim + 127) / 128; p.n_tiles = 1;
  p.kblocks_total = (M + 63) / 64;
  p.kblocks_per_split = (p.kblocks_total + splits - 1) / splits;
  p.splits = (p.kblocks_total + p.kblocks_per_split - 1) / p.kblocks_per_split;
  p.out = (float*)workspace; p.ldo = kTokLd; p.split_stride = (long long)dim * kTokLd;
  cudaStream_t st = (cudaStream_t)stream;
  rc = launch_gemm<true>(ta, tb, p, st);
  if (rc) return rc;
  const int total = dim * 193;
  launch_pdl((embed_dw_unpermute_kernel), dim3((total + 255) / 256), dim3(256), (size_t)(0), st, (const float*)workspace, p.splits, dim, grad_w, grad_b,
                                                                  accumulate ? 1 : 0);
  return check_launch("lafs_embed_bwd_weight_perm");
}

extern "C" int lafs_embed_bwd_weight(const void* grad_emb_bf16, const void* tokens_bf16, int M, int dim, float* grad_w,
                                     void* workspace, size_t workspace_bytes, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(grad_emb_bf16)) return brc;
  LAFS_REQUIRE(grad_emb_bf16 && tokens_bf16 && grad_w && workspace, LAFS_ERR_ARG, "lafs_embed_bwd_weight: null pointer");
  LAFS_REQUIRE(M > 0 && dim > 0 && dim % 8 == 0, LAFS_ERR_ARG, "lafs_embed_bwd_weight: M=%d dim=%d (dim must be a multiple of 8)", M, dim);
  LAFS_REQUIRE((((uintptr_t)grad_emb_bf16 | (uintptr_t)tokens_bf16 | (uintptr_t)grad_w | (uintptr_t)workspace) & 15u) == 0,
               LAFS_ERR_ARG, "lafs_embed_bwd_weight: pointers must be 16-byte aligned");
  const int splits = embed_bwd_splits(M, dim);
  const size_t need = (size_t)splits * dim * 192 * sizeof(float);
  LAFS_REQUIRE(workspace_bytes >= need, LAFS_ERR_WORKSPACE, "lafs_embed_bwd_weight: workspace %zu < %zu", workspace_bytes, need);
  CUtensorMap ta, tb;
  int rc = encode_bf16_2d(&ta, grad_emb_bf16, (uint64_t)M, (uint64_t)dim, (uint64_t)dim * 2, 64, 64);   // MN-major A: [K=M, dim]
  if (rc) return rc;
  rc = encode_bf16_2d(&tb, tokens_bf16, (uint64_t)M, 192, 192 * 2, 64, 64);                             // MN-major B: [K=M, 192]
  if (rc) return rc;
  GemmParams p{};
  p.M = dim; p.N = 192; p.K = M;
  p.m_tiles = (dim + 127) / 128; p.n_tiles = 1;
  p.kblocks_total = (M + 63) / 64;
  p.kblocks_per_split = (p.kblocks_total + splits - 1) / splits;
  p.splits = (p.kblocks_total + p.kblocks_per_split - 1) / p.kblocks_per_split;
  p.out = (float*)workspace; p.ldo = 192; p.split_stride = (long long)dim * 192;
  cudaStream_t st = (cudaStream_t)stream;
  rc = launch_gemm<true>(ta, tb, p, st);
  if (rc) return rc;
  const long long n = (long long)dim * 192;
  launch_pdl((split_reduce_kernel), dim3((unsigned)((n + 255) / 256)), dim3(256), (size_t)(0), st, (const float*)workspace, p.splits, n, n, grad_w);
  return check_launch("lafs_embed_bwd_weight");
}

extern "C" int lafs_embed_bwd_tokens(const void* grad_emb_bf16, const void* weight_bf16, int M, int dim, float* grad_tokens,
                                     lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(grad_emb_bf16)) return brc;
  LAFS_REQUIRE(grad_emb_bf16 && weight_bf16 && grad_tokens, LAFS_ERR_ARG, "lafs_embed_bwd_tokens: null pointer");
  LAFS_REQUIRE(M > 0 && dim > 0 && dim % 8 == 0, LAFS_ERR_ARG, "lafs_embed_bwd_tokens: M=%d dim=%d (dim must be a multiple of 8)", M, dim);
  LAFS_REQUIRE((((uintptr_t)grad_emb_bf16 | (uintptr_t)weight_bf16 | (uintptr_t)grad_tokens) & 15u) == 0, LAFS_ERR_ARG,
               "lafs_embed_bwd_tokens: pointers must be 16-byte aligned");
  CUtensorMap ta, tb;
  int rc = encode_bf16_2d(&ta, grad_emb_bf16, (uint64_t)M, (uint64_t)dim, (uint64_t)dim * 2, 64, 128);  // K-major A: [M, K=dim]
  if (rc) return rc;
  rc = encode_bf16_2d(&tb, weight_bf16, (uint64_t)dim, 192, 192 * 2, 64, 64);                           // MN-major B: [K=dim, 192]
  if (rc) return rc;
  GemmParams p{};
  p.M = M; p.N = 192; p.K = dim;
  p.m_tiles = (M + 127) / 128; p.n_tiles = 1; p.splits = 1;
  p.kblocks_total = (dim + 63) / 64; p.kblocks_per_split = p.kblocks_total;
  p.out = grad_tokens; p.ldo = 192; p.split_stride = 0;
  return launch_gemm<false>(ta, tb, p, (cudaStream_t)stream);
}

/* lafs_head_bwd_weight with the Jacobian's rank-one term computed on
 * the tensor core from the per-class dots of lafs_head_grad_logits_t (tpart [tparts][ldt], tparts = 4*ceil(B/128)):
 * one GEMM, no normalize_bwd pass. */
extern "C" int lafs_head_bwd_weight_t(const void* grad_bf16, long long ldg, const void* e_hat, const void* w_hat,
                                      const float* inv_norm_w, float* tpart, int tparts, long long ldt, int B,
                                      int C_local, int D, float* grad_w, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(grad_bf16)) return brc;
  LAFS_REQUIRE(grad_bf16 && e_hat && w_hat && inv_norm_w && tpart && grad_w, LAFS_ERR_ARG, "lafs_head_bwd_weight_t: null pointer");
  LAFS_REQUIRE(B > 0 && C_local > 0 && D > 0 && D % 64 == 0 && D <= 768, LAFS_ERR_ARG, "lafs_head_bwd_weight_t: B=%d C=%d D=%d", B, C_local, D);
  LAFS_REQUIRE(ldg % 8 == 0 && ldg >= C_local, LAFS_ERR_ARG, "lafs_head_bwd_weight_t: ldg=%lld must be a multiple of 8 and >= C_local", ldg);
  LAFS_REQUIRE(tparts > 0 && ldt >= C_local, LAFS_ERR_ARG, "lafs_head_bwd_weight_t: tparts=%d ldt=%lld", tparts, ldt);
  CUtensorMap ta, tb, tw;
  int rc = encode_bf16_2d(&ta, grad_bf16, (uint64_t)B, (uint64_t)C_local, (uint64_t)ldg * 2, 64, 64);  // MN-major A: [K=B, M=C]
  if (rc) return rc;
  rc = encode_bf16_2d(&tb, e_hat, (uint64_t)B, (uint64_t)D, (uint64_t)D * 2, 64, 64);                  // MN-major B: [K=B, N=D]
  if (rc) return rc;
  rc = encode_bf16_2d(&tw, w_hat, (uint64_t)C_local, (uint64_t)D, (uint64_t)D * 2, 64, 64);            // MN-major B: [K=classes, N=D]
  if (rc) return rc;
  GemmParams p{};
  p.M = C_local; p.N = D; p.K = B;
  p.m_tiles = (C_local + 127) / 128; p.n_tiles = (D + 255) / 256; p.splits = 1;
  p.kblocks_total = (B + 63) / 64; p.kblocks_per_split = p.kblocks_total;
  p.out = grad_w; p.ldo = D; p.split_stride = 0;
  cudaStream_t st = (cudaStream_t)stream;
  launch_pdl((t_reduce_kernel), dim3((C_local + 255) / 256), dim3(256), (size_t)(0), st, tpart, tparts, ldt, C_local);   // row 0 <- sum of the rows
  DiagParams dp{tpart, inv_norm_w};
  cudaError_t e = cudaFuncSetAttribute(dw_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, gb::kSmem);
  LAFS_REQUIRE(e == cudaSuccess, LAFS_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int total = p.m_tiles * p.n_tiles;
  launch_pdl((dw_diag_kernel), dim3(total < kNumSMs ? total : kNumSMs), dim3(gb::kThreads), (size_t)(gb::kSmem), st, ta, tb, tw, p, dp);
  return check_launch("lafs_head_bwd_weight_t");
}

/* out [M, N] fp32 = A^T . B  with A [Kr rows, M cols] and B [Kr rows, N cols] bf16 row-major (both operands MN-major;
 * the contraction runs over the rows): the dW GEMM above without a Jacobian pass.  Used by the fused DINO head
 * (dino_head.cu): dW_raw [K, D] = (P_s ; Q)^T . (cnt X_hat_s ; -X~). */
extern "C" int lafs_gemm_tn(const void* a_bf16, long long lda, const void* b_bf16, long long ldb, int Kr, int M, int N,
                            float* out, long long ldo, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(a_bf16)) return brc;
  LAFS_REQUIRE(a_bf16 && b_bf16 && out, LAFS_ERR_ARG, "lafs_gemm_tn: null pointer");
  LAFS_REQUIRE(Kr > 0 && M > 0 && N > 0, LAFS_ERR_ARG, "lafs_gemm_tn: Kr=%d M=%d N=%d", Kr, M, N);
  LAFS_REQUIRE(lda % 8 == 0 && lda >= M && ldb % 8 == 0 && ldb >= N && ldo % 4 == 0 && ldo >= N, LAFS_ERR_ARG,
               "lafs_gemm_tn: lda=%lld ldb=%lld (multiples of 8, >= M / N), ldo=%lld (multiple of 4, >= N)", lda, ldb, ldo);
  LAFS_REQUIRE((((uintptr_t)a_bf16 | (uintptr_t)b_bf16 | (uintptr_t)out) & 15u) == 0, LAFS_ERR_ARG,
               "lafs_gemm_tn: pointers must be 16-byte aligned");
  CUtensorMap ta, tb;
  int rc = encode_bf16_2d(&ta, a_bf16, (uint64_t)Kr, (uint64_t)M, (uint64_t)lda * 2, 64, 64);   // MN-major A: [K=Kr, M]
  if (rc) return rc;
  rc = encode_bf16_2d(&tb, b_bf16, (uint64_t)Kr, (uint64_t)N, (uint64_t)ldb * 2, 64, 64);       // MN-major B: [K=Kr, N]
  if (rc) return rc;
  GemmParams p{};
  p.M = M; p.N = N; p.K = Kr;
  p.m_tiles = (M + 127) / 128; p.n_tiles = (N + 255) / 256; p.splits = 1;
  p.kblocks_total = (Kr + 63) / 64; p.kblocks_per_split = p.kblocks_total;
  p.out = out; p.ldo = ldo; p.split_stride = 0;
  return launch_gemm<true>(ta, tb, p, (cudaStream_t)stream);
}
