// (4) CosFace / ArcFace margin head: normalised-embedding x normalised-weight GEMM on tcgen05
// tensor cores with the margin, online softmax and cross-entropy statistics fused into the
// TMEM epilogue -- the [B, C] logits are never materialised on the loss path.
// Replaces CosFace.forward (face_pre_pro/ViT_face.py:49-89: CPU one_hot + H2D, two normalize
// passes, ~8 element-wise passes over [B,C] fp32) + CrossEntropyLoss / SoftTargetCrossEntropy
// (train_largescale.py:601-604,820).
//
// GEMM orientation: M = batch rows (128 per CTA, the E tile stays resident in shared memory),
// N = classes (BN per UMMA), K = D.  One TMEM lane = one batch row, so the per-row softmax
// statistics are plain sequential reductions over the columns a thread reads back (no shuffles).
// Warp roles: 0 = TMA producer, 1 = UMMA issuer (+TMEM owner), 2..5 = epilogue.
// Classes are sharded across GPUs with torch.chunk's rule (ViT_face.py:56); a rank passes its
// shard [class_lo, class_lo + C_local) and gets per-row (max, sum-exp, target-logit) partials.
#include <stdlib.h>
#include "umma.cuh"
#include "../../include/lafs_b200.h"

namespace lafs {

using namespace umma;

constexpr int kHeadThreads = 320;   // warp 0 TMA, warp 1 UMMA issue + TMEM owner, warps 2..9 epilogue
constexpr int kEpiWarps = 8;        // two warps per TMEM lane quarter, each takes half of a tile's columns
constexpr int kBM = 128;
constexpr float kLog2eH = 1.4426950408889634f;
constexpr float kLn2H = 0.6931471805599453f;

enum : int { HEAD_STATS = 0, HEAD_LOGITS = 1, HEAD_GRAD = 2, HEAD_GRAD_T = 3 };   // GRAD_T: GRAD + per-class column dots

struct HeadParams {
  int B, C_local, D, class_lo;       // class_lo: global id of local class 0
  int kch;                           // D / 64
  int stages;                        // B-operand ring depth
  int nranges;                       // CTAs per M tile
  int nchunks;                       // ceil(C_local / BN)
  float s, m, lam;                   // scale, margin, mixup lambda (1 for hard labels)
  int kind;                          // 0 CosFace, 1 ArcFace
  float cos_m, sin_m, th, mm;        // ArcFace constants
  const int64_t* label_a;            // [B] global class ids
  const int64_t* label_b;            // [B] or nullptr
  float* part;                       // STATS: [B][2*nranges][4]
  float* logits;                     // LOGITS: [B][ldc] fp32
  long long ldc;
  const float* row_lse2;             // GRAD: [B] log2-domain lse of the full row
  __nv_bfloat16* grad;               // GRAD: [B][ldg] bf16  (softmax - target) * gscale
  long long ldg;
  int g256;                          // GRAD: rows are 32-byte aligned -> 256-bit stores
  float gscale;                      // s / B_global
  const float* grad_out;             // device scalar: upstream gradient of the loss
  int debug;                         // development ablation (env LAFS_HEAD_DEBUG): 1 = GRAD modes skip the G stores
  // GRAD_T reuses `part` (= tpart [mtiles*4][ldt]: partial column dots sum_b G[b,c]*cos[b,c]) and `ldc` (= ldt),
  // so that the parameter block -- and with it the code of the measured kernels -- stays exactly as it was
};

// margin-adjusted cosine (logit / s) of a class whose target weight is t (0 < t <= 1)
__device__ __forceinline__ float margin_cos(float cosv, float t, const HeadParams& p) {
  if (p.kind == 0) return cosv - p.m * t;  // CosFace, soft targets: s*(cos - m*t)
  const float sine = sqrtf(fminf(fmaxf(1.f - cosv * cosv, 0.f), 1.f));
  const float phi = cosv * p.cos_m - sine * p.sin_m;
  return cosv > p.th ? phi : cosv - p.mm;  // ArcFace (hard labels only)
}

// d(margin_logit)/d(cos) / s for the target class (1 for CosFace and for non-target classes)
__device__ __forceinline__ float margin_dcos(float cosv, const HeadParams& p) {
  if (p.kind == 0) return 1.f;
  if (!(cosv > p.th)) return 1.f;
  const float sine = sqrtf(fminf(fmaxf(1.f - cosv * cosv, 1e-12f), 1.f));
  return p.cos_m + cosv * p.sin_m / sine;
}

// PAIR: two CTAs (one cluster, one TPC) hold two consecutive M tiles and HALF of every W chunk each;
// the leader issues cta_group::2 UMMAs of M = 256 that feed both CTAs' TMEM.  Per SM this halves the
// L2 -> shared-memory fill and the shared-memory operand reads of the streamed W operand, which is
// what bounds the single-CTA kernel (E stays resident, so W is the only stream).
template <int BN, int MODE, bool PAIR>
__global__ void __launch_bounds__(kHeadThreads, 1)
head_gemm_kernel(const __grid_constant__ CUtensorMap tmap_e, const __grid_constant__ CUtensorMap tmap_w,
                 const HeadParams p) {
  pdl_wait();
  constexpr int kStageRows = PAIR ? BN / 2 : BN;     // W rows this CTA loads per k step
  constexpr int kStageBytes = kStageRows * 128;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [E tile: kch x 16 KB][W ring: stages x kStageBytes][barriers]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_e = smem;
  uint8_t* smem_b = smem_e + (size_t)p.kch * (kBM * 128);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.stages * kStageBytes);
  uint64_t* e_full = bars;                      // 1   (PAIR: the leader's is used)
  uint64_t* b_full = bars + 1;                  // stages (PAIR: the leader's are used)
  uint64_t* b_empty = b_full + p.stages;        // stages (own; PAIR: signalled by the leader's multicast commit)
  uint64_t* acc_full = b_empty + p.stages;      // 2      (own; same)
  uint64_t* acc_empty = acc_full + 2;           // 2      (PAIR: the leader's, both CTAs' epilogues arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, range = blockIdx.y;
  const int row0 = mt * kBM;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_e);
    prefetch_tensormap(&tmap_w);
    mbar_init(e_full, 1);
    for (int i = 0; i < p.stages; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, kEpiWarps * (PAIR ? 2 : 1)); }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_pair(tmem_slot, 2 * BN);
    else tmem_alloc(tmem_slot, 2 * BN);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();      // peer barriers initialised before any remote arrive / multicast
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // chunks handled by this CTA (pair): range, range + nranges, ...
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint32_t e_bar = PAIR ? mapa_shared(smem_u32(e_full), 0) : 0u;
      if (leader) mbar_arrive_expect_tx(e_full, (uint32_t)p.kch * (kBM * 128) * (PAIR ? 2u : 1u));
      for (int k = 0; k < p.kch; ++k) {
        if (PAIR) tma_load_2d_pair(smem_e + (size_t)k * (kBM * 128), &tmap_e, e_bar, k * 64, row0);
        else tma_load_2d(smem_e + (size_t)k * (kBM * 128), &tmap_e, e_full, k * 64, row0);
      }
      int stage = 0;
      uint32_t phase = 0;
      // W is read once from HBM (plus L2 hits by the other M tiles): the ring alone (<= 96 KB per SM)
      // cannot cover HBM latency at the UMMA consumption rate (ncu: the epilogue warps wait on acc_full
      // for 25 % of the samples), so the producer also pulls the chunk after next into L2
      const int pf_ahead = 2 * p.nranges;
      const bool pf_owner = PAIR ? blockIdx.x < 2 : blockIdx.x == 0;   // one prefetch per W tile, not one per M tile
      for (int chunk = range; chunk < p.nchunks; chunk += p.nranges) {
        const int wrow = chunk * BN + (int)rank * kStageRows;
        const int pf_row = wrow + pf_ahead * BN;
        const bool pf = pf_owner && chunk + pf_ahead < p.nchunks;
        for (int k = 0; k < p.kch; ++k) {
          if (pf) tma_prefetch_l2_2d(&tmap_w, k * 64, pf_row);
          mbar_wait(b_empty + stage, phase ^ 1);     // every consumer of this slot (both CTAs' data) is done
          if (leader) mbar_arrive_expect_tx(b_full + stage, (uint32_t)kStageBytes * (PAIR ? 2u : 1u));
          uint8_t* dst = smem_b + (size_t)stage * kStageBytes;
          if (PAIR) tma_load_2d_pair(dst, &tmap_w, mapa_shared(smem_u32(b_full + stage), 0), k * 64, wrow);
          else tma_load_2d(dst, &tmap_w, b_full + stage, k * 64, wrow);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== UMMA issuer (PAIR: the leader only) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 2 * kBM : kBM, BN);
      mbar_wait(e_full, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int chunk = range; chunk < p.nchunks; chunk += p.nranges, ++it) {
        const int buf = it & 1;
        const uint32_t use = (uint32_t)(it >> 1);
        mbar_wait(acc_empty + buf, (use & 1) ^ 1);   // epilogue(s) have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
        for (int k = 0; k < p.kch; ++k) {
          mbar_wait(b_full + stage, phase);
          tc_fence_after();
          const uint64_t da = make_desc_k_sw128(smem_u32(smem_e + (size_t)k * (kBM * 128)));
          const uint64_t db = make_desc_k_sw128(smem_u32(smem_b + (size_t)stage * kStageBytes));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (PAIR) mma_f16_ss_pair(d_tmem, desc_advance_k(da, kk * 16), desc_advance_k(db, kk * 16), idesc, (k | kk) != 0);
            else mma_f16_ss(d_tmem, desc_advance_k(da, kk * 16), desc_advance_k(db, kk * 16), idesc, (k | kk) != 0);
          }
          // frees the smem stage (in both CTAs) when these UMMAs finish
          if (PAIR) mma_commit_pair(b_empty + stage, 3);
          else mma_commit(b_empty + stage);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue(s)
        if (PAIR) mma_commit_pair(acc_full + buf, 3);
        else mma_commit(acc_full + buf);
      }
    }
  } else {
    // ===================== epilogue warps (TMEM lane quarter = warp % 4) =====================
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;          // which half of the tile's columns
    const int row = quarter * 32 + lane;
    const int b = row0 + row;
    const bool row_ok = b < p.B;
    long long la = -1, lb = -1;
    float ta = 1.f, tb = 0.f;   // target weights of label_a / label_b
    if (row_ok) {
      la = p.label_a[b] - p.class_lo;
      if (p.label_b != nullptr) {
        lb = p.label_b[b] - p.class_lo;
        ta = p.lam; tb = 1.f - p.lam;
        if (lb == la) { ta = 1.f; tb = 0.f; lb = -1; }     // same class twice: weights add up
      }
    }
    const float sk2 = p.s * kLog2eH;         // cosine -> log2-domain logit
    float m_run = -INFINITY, l_run = 0.f, tgt_a = 0.f, tgt_b = 0.f;
    float lse2 = 0.f;
    float gscale = 0.f;
    if ((MODE == HEAD_GRAD || MODE == HEAD_GRAD_T) && row_ok) { lse2 = p.row_lse2[b]; gscale = p.gscale * __ldg(p.grad_out); }
    const uint32_t acc_empty_remote = (PAIR && !leader) ? mapa_shared(smem_u32(acc_empty), 0) : 0u;
    int it = 0;
    for (int chunk = range; chunk < p.nchunks; chunk += p.nranges, ++it) {
      const int buf = it & 1;
      const uint32_t use = (uint32_t)(it >> 1);
      mbar_wait(acc_full + buf, use & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(buf * BN) + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
      for (int piece = half * (BN / 64); piece < (half + 1) * (BN / 64); ++piece) {
        uint32_t raw[32];
        tmem_ld_32x32b_x32(taddr + (uint32_t)(piece * 32), raw);
        tmem_ld_wait();
        const int cbase = chunk * BN + piece * 32;            // local class id of column 0
        if (cbase >= p.C_local) break;                        // fully out-of-range piece
        float c[32];                                          // margin-adjusted cosines
#pragma unroll
        for (int j = 0; j < 32; ++j) c[j] = __uint_as_float(raw[j]);
        // rare paths: target column(s) inside this piece, or ragged tail of the shard
        const bool has_a = (unsigned long long)(la - cbase) < 32ull;
        const bool has_b = (unsigned long long)(lb - cbase) < 32ull;
        const bool tail = cbase + 32 > p.C_local;
        if (has_a || has_b || tail) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int cc = cbase + j;
            if (has_a && cc == (int)la) { c[j] = margin_cos(__uint_as_float(raw[j]), ta, p); tgt_a = p.s * c[j]; }
            if (has_b && cc == (int)lb) { c[j] = margin_cos(__uint_as_float(raw[j]), tb, p); tgt_b = p.s * c[j]; }
            if (cc >= p.C_local) c[j] = -INFINITY;
          }
        }
        if (MODE == HEAD_STATS) {
          float pm[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            pm[q] = c[q];
#pragma unroll
            for (int j = 4 + q; j < 32; j += 4) pm[q] = fmaxf(pm[q], c[j]);
          }
          const float m_new = fmaxf(m_run, fmaxf(fmaxf(pm[0], pm[1]), fmaxf(pm[2], pm[3])) * sk2);
          l_run *= ex2(m_run - m_new);
          float acc[4] = {0.f, 0.f, 0.f, 0.f};                // independent chains: latency, not issue, bound
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[j & 3] += ex2(fmaf(c[j], sk2, -m_new));
          l_run += (acc[0] + acc[1]) + (acc[2] + acc[3]);
          m_run = m_new;
        } else if (MODE == HEAD_LOGITS) {
          if (row_ok) {
            float* dst = p.logits + (long long)b * p.ldc + cbase;
            if (!tail && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(dst + j) = make_float4(p.s * c[j], p.s * c[j + 1], p.s * c[j + 2], p.s * c[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (cbase + j < p.C_local) dst[j] = p.s * c[j];
            }
          }
        } else if (MODE == HEAD_GRAD_T) {
          // HEAD_GRAD plus t[c] = sum_b G[b,c] * cos[b,c] (= <w_hat_c, dW_hat_c>) for the F.normalize Jacobian of
          // dW: each warp reduces its 32 rows per column with a transposing butterfly (lane L ends with
          // column L) and stores one coalesced 128-byte partial per piece; rows past B carry gscale = 0.
          float gv[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) gv[j] = ex2(fmaf(c[j], sk2, -lse2));
          if (has_a || has_b) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int cc = cbase + j;
              if (has_a && cc == (int)la) gv[j] = (gv[j] - ta) * margin_dcos(__uint_as_float(raw[j]), p);
              if (has_b && cc == (int)lb) gv[j] -= tb;
            }
          }
          uint32_t pk32[16];
          float pr[32];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            pk32[j >> 1] = Half2Ops<__nv_bfloat16>::pack(gv[j] * gscale, gv[j + 1] * gscale);
            // the bf16 values the dW GEMM will consume, times the raw cosine
            pr[j] = Half2Ops<__nv_bfloat16>::lo(pk32[j >> 1]) * __uint_as_float(raw[j]);
            pr[j + 1] = Half2Ops<__nv_bfloat16>::hi(pk32[j >> 1]) * __uint_as_float(raw[j + 1]);
          }
          if (row_ok && !(p.debug & 1)) {
            __nv_bfloat16* dst = p.grad + (long long)b * p.ldg + cbase;
            if (!tail) {
              if (p.g256) {
                st_global_256(dst, make_uint4(pk32[0], pk32[1], pk32[2], pk32[3]), make_uint4(pk32[4], pk32[5], pk32[6], pk32[7]));
                st_global_256(dst + 16, make_uint4(pk32[8], pk32[9], pk32[10], pk32[11]),
                              make_uint4(pk32[12], pk32[13], pk32[14], pk32[15]));
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  *reinterpret_cast<uint4*>(dst + 8 * j) = make_uint4(pk32[4 * j], pk32[4 * j + 1], pk32[4 * j + 2], pk32[4 * j + 3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (cbase + j < p.C_local) dst[j] = __float2bfloat16_rn(gv[j] * gscale);
            }
          }
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int i = 0; i < o; ++i) {
              const float send = up ? pr[i] : pr[i + o];
              const float keep = up ? pr[i + o] : pr[i];
              pr[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
          p.part[(size_t)(mt * 4 + quarter) * (size_t)p.ldc + cbase + lane] = pr[0];   // part = tpart, ldc = ldt >= round32(C_local)
        } else {  // HEAD_GRAD: (softmax - target) * dz/dcos * gscale, bf16; 64 contiguous bytes per row
          if (row_ok) {
            float gv[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) gv[j] = ex2(fmaf(c[j], sk2, -lse2));
            if (has_a || has_b) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int cc = cbase + j;
                if (has_a && cc == (int)la) gv[j] = (gv[j] - ta) * margin_dcos(__uint_as_float(raw[j]), p);
                if (has_b && cc == (int)lb) gv[j] -= tb;
              }
            }
            __nv_bfloat16* dst = p.grad + (long long)b * p.ldg + cbase;   // ldg % 8 == 0: 16-byte aligned
            if (p.debug & 1) {
              if (gv[5] == 123.456f) dst[0] = __float2bfloat16_rn(gv[7]);   // keeps the arithmetic alive, never true
            } else if (!tail) {
              uint4 pk[4];
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                pk[j >> 3].x = Half2Ops<__nv_bfloat16>::pack(gv[j] * gscale, gv[j + 1] * gscale);
                pk[j >> 3].y = Half2Ops<__nv_bfloat16>::pack(gv[j + 2] * gscale, gv[j + 3] * gscale);
                pk[j >> 3].z = Half2Ops<__nv_bfloat16>::pack(gv[j + 4] * gscale, gv[j + 5] * gscale);
                pk[j >> 3].w = Half2Ops<__nv_bfloat16>::pack(gv[j + 6] * gscale, gv[j + 7] * gscale);
              }
              if (p.g256) {          // two full 32-byte sectors per lane instead of four half sectors
                st_global_256(dst, pk[0], pk[1]);
                st_global_256(dst + 16, pk[2], pk[3]);
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + 8 * j) = pk[j];
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (cbase + j < p.C_local) dst[j] = __float2bfloat16_rn(gv[j] * gscale);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && !leader) mbar_arrive_remote_relaxed(acc_empty_remote + (uint32_t)(buf * 8));
        else mbar_arrive_relaxed(acc_empty + buf);
      }
    }
    if (MODE == HEAD_STATS && row_ok) {
      float4 o = make_float4(m_run, l_run, tgt_a, tgt_b);
      *reinterpret_cast<float4*>(p.part + ((size_t)b * (2 * p.nranges) + (range * 2 + half)) * 4) = o;
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();      // the peer's smem / barriers stay valid until both CTAs are done
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, 2 * BN);
    else tmem_dealloc(tmem_base, 2 * BN);
  }
}

// merge `nparts` partial (max2, sumexp, tgt_a, tgt_b) records per row -> one record per row.
// One warp per row, lanes stride over the parts, butterfly merge.
__global__ void __launch_bounds__(256)
head_merge_kernel(const float* __restrict__ part, int B, int nparts, long long part_stride,
                  long long row_stride, float* __restrict__ out) {
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;
  float m = -INFINITY, l = 0.f, ta = 0.f, tb = 0.f;
  for (int r = lane; r < nparts; r += 32) {
    const float4 v = *reinterpret_cast<const float4*>(part + (size_t)r * part_stride + (size_t)b * row_stride);
    if (v.x == -INFINITY) continue;     // CTA saw no valid class for this row
    const float mn = fmaxf(m, v.x);
    l = l * ex2(m - mn) + v.y * ex2(v.x - mn);
    m = mn;
    ta += v.z;
    tb += v.w;
  }
  const float mw = warp_max(m);
  l = warp_sum(m == -INFINITY ? 0.f : l * ex2(m - mw));
  ta = warp_sum(ta);
  tb = warp_sum(tb);
  if (lane == 0) *reinterpret_cast<float4*>(out + (size_t)b * 4) = make_float4(mw, l, ta, tb);
}

// per-row loss from merged statistics: lse(z) - (ta_w*z_a + tb_w*z_b); mean over rows.
__global__ void __launch_bounds__(256)
head_loss_kernel(const float* __restrict__ stats, const int64_t* __restrict__ label_a,
                 const int64_t* __restrict__ label_b, float lam, int B, float* __restrict__ row_lse2,
                 float* __restrict__ loss_out) {
  pdl_wait();
  __shared__ float red[256];
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += 256) {
    const float4 v = *reinterpret_cast<const float4*>(stats + (size_t)b * 4);
    const float lse2 = v.x + lg2(v.y);
    row_lse2[b] = lse2;
    float wa = 1.f, wb = 0.f;
    if (label_b != nullptr && label_b[b] != label_a[b]) { wa = lam; wb = 1.f - lam; }
    acc += kLn2H * lse2 - (wa * v.z + wb * v.w);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss_out = red[0] / (float)B;
}

// ---- operand preparation --------------------------------------------------------------------
// rows of x [R, D] -> L2-normalised bf16 rows (+ 1/max(||x||, eps), F.normalize's eps=1e-12)
template <typename T>
__global__ void __launch_bounds__(256)
normalize_rows_kernel(const T* __restrict__ x, int R, int D, __nv_bfloat16* __restrict__ out,
                      float* __restrict__ inv_norm) {
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (r >= R) return;
  const T* src = x + (size_t)r * D;
  if constexpr (sizeof(T) == 4) {
    // fp32 master weights (the [C, D] pass of every step): a lane keeps its <= 6 x float4 of the row
    // in registers, so the row is read ONCE with 128-bit streaming loads and written as 64-bit bf16 quads
    if (D % 128 == 0 && D <= 768 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0) {
      float4 v[6];
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        if (i * 128 < D) {
          v[i] = ld_stream_f4(src + i * 128 + lane * 4);
          ss = fmaf(v[i].x, v[i].x, ss); ss = fmaf(v[i].y, v[i].y, ss);
          ss = fmaf(v[i].z, v[i].z, ss); ss = fmaf(v[i].w, v[i].w, ss);
        }
      }
      ss = warp_sum(ss);
      const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
      if (lane == 0 && inv_norm != nullptr) inv_norm[r] = inv;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        if (i * 128 < D) {
          uint2 o;
          o.x = Half2Ops<__nv_bfloat16>::pack(v[i].x * inv, v[i].y * inv);
          o.y = Half2Ops<__nv_bfloat16>::pack(v[i].z * inv, v[i].w * inv);
          *reinterpret_cast<uint2*>(out + (size_t)r * D + i * 128 + lane * 4) = o;
        }
      }
      return;
    }
  }
  float ss = 0.f;
  for (int i = lane; i < D; i += 32) {
    const float v = (float)src[i];
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  if (lane == 0 && inv_norm != nullptr) inv_norm[r] = inv;
  for (int i = lane; i < D; i += 32) out[(size_t)r * D + i] = __float2bfloat16_rn((float)src[i] * inv);
}

TmaEncoder::EncodeTiled TmaEncoder::get() {
  static EncodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiled>(p);
  }
  return fn;
}

int TmaEncoder::bf16_2d_sw128(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                              uint64_t row_stride_bytes, uint32_t box_rows) {
  EncodeTiled enc = get();
  LAFS_REQUIRE(enc != nullptr, LAFS_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LAFS_REQUIRE(r == CUDA_SUCCESS, LAFS_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu", (int)r,
               (unsigned long long)rows, (unsigned long long)cols);
  return LAFS_OK;
}

struct HeadLaunch {
  int BN, stages, nranges, nchunks, mtiles;
  bool pair;
  size_t smem;
};

// shape-only part of the plan (no device queries): tile width and the 1-CTA ring depth
static int plan_head(int B, int C_local, int D, HeadLaunch* hl) {
  LAFS_REQUIRE(D % 64 == 0 && D >= 64 && D <= 768, LAFS_ERR_ARG, "margin head: D=%d must be a multiple of 64 in [64,768]", D);
  const int kch = D / 64;
  const size_t e_bytes = (size_t)kch * kBM * 128;
  const size_t budget = 227 * 1024 - 2048;   // barriers + 1 KB alignment slack
  int BN = 256;
  if (e_bytes + 2 * (size_t)BN * 128 > budget) BN = 128;
  int stages = (int)((budget - e_bytes) / ((size_t)BN * 128));
  if (stages > 6) stages = 6;
  LAFS_REQUIRE(stages >= 2, LAFS_ERR_ARG, "margin head: D=%d leaves no room for a 2-stage pipeline", D);
  hl->BN = BN;
  hl->stages = stages;
  hl->pair = false;
  hl->mtiles = (B + kBM - 1) / kBM;
  hl->nchunks = (C_local + BN - 1) / BN;
  int per = kNumSMs / hl->mtiles;
  if (per < 1) per = 1;
  hl->nranges = hl->nchunks < per ? hl->nchunks : per;
  hl->smem = e_bytes + (size_t)stages * BN * 128 + 1024 /*align*/ + 256 /*barriers*/;
  return LAFS_OK;
}

template <int BN, int MODE, bool PAIR>
static int launch_head(const CUtensorMap& te, const CUtensorMap& tw, const HeadParams& p, const HeadLaunch& hl,
                       cudaStream_t st) {
  auto kern = head_gemm_kernel<BN, MODE, PAIR>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hl.smem);
  LAFS_REQUIRE(e == cudaSuccess, LAFS_ERR_CUDA, "cudaFuncSetAttribute(smem=%zu): %s", hl.smem, cudaGetErrorString(e));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(hl.mtiles, hl.nranges);
  cfg.blockDim = dim3(kHeadThreads);
  cfg.dynamicSmemBytes = hl.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = PAIR ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // see common.cuh: every kernel starts with pdl_wait()
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  e = cudaLaunchKernelEx(&cfg, kern, te, tw, p);
  LAFS_REQUIRE(e == cudaSuccess, LAFS_ERR_CUDA, "head_gemm_kernel launch: %s", cudaGetErrorString(e));
  return check_launch("head_gemm_kernel");
}

// CTA-pair variant of the plan: possible when the M tiles pair up and a 256-class chunk fits; each
// CTA then stages only half a chunk (16 KB) per k step, so the ring gets deeper.  The number of
// co-resident pairs is asked from the driver (pairs must sit on one TPC).
template <int MODE>
static void plan_pair(HeadLaunch* hl, int C_local, int D) {
  if (hl->mtiles % 2 != 0) return;
  if (const char* e = getenv("LAFS_HEAD_1SM")) { if (atoi(e) != 0) return; }
  const size_t e_bytes = (size_t)(D / 64) * kBM * 128;
  const size_t budget = 227 * 1024 - 2048;
  const size_t stage = 128 * 128;            // half of a 256-class chunk, 64 k
  if (e_bytes + 2 * stage > budget) return;
  int stages = (int)((budget - e_bytes) / stage);
  if (stages > 8) stages = 8;
  const size_t smem = e_bytes + (size_t)stages * stage + 1024 + 256;
  auto kern = head_gemm_kernel<256, MODE, true>;
  static int cached = -1;                     // co-resident pairs for (MODE, smem): queried once per process
  static size_t cached_smem = 0;
  if (cached < 0 || cached_smem != smem) {
    int n = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2, 1);
      cfg.blockDim = dim3(kHeadThreads);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) n = 0;
    }
    cudaGetLastError();
    cached = n; cached_smem = smem;
  }
  const int nclusters = cached;
  if (nclusters < 1) return;
  int ctas = 2 * nclusters;
  if (ctas > kNumSMs) ctas = kNumSMs;
  const int nchunks = (C_local + 255) / 256;
  int per = ctas / hl->mtiles;
  if (per < 1) return;
  hl->pair = true;
  hl->BN = 256;
  hl->stages = stages;
  hl->smem = smem;
  hl->nchunks = nchunks;
  hl->nranges = nchunks < per ? nchunks : per;
}

template <int MODE>
static int launch_head_any(const CUtensorMap& te, const CUtensorMap& tw, const HeadParams& p, const HeadLaunch& hl,
                           cudaStream_t st) {
  if (hl.pair) return launch_head<256, MODE, true>(te, tw, p, hl, st);
  return hl.BN == 256 ? launch_head<256, MODE, false>(te, tw, p, hl, st) : launch_head<128, MODE, false>(te, tw, p, hl, st);
}

template <int MODE>
static int head_common(const void* e_hat, const void* w_hat, const int64_t* label_a, const int64_t* label_b, float lam,
                       int B, int C_local, int D, int class_lo, float s, float m, int kind, HeadParams* p,
                       HeadLaunch* hl, CUtensorMap* te, CUtensorMap* tw, const char* who) {
  LAFS_REQUIRE(e_hat && w_hat && label_a, LAFS_ERR_ARG, "%s: null pointer", who);
  LAFS_REQUIRE(B > 0 && C_local > 0, LAFS_ERR_ARG, "%s: B=%d C_local=%d", who, B, C_local);
  LAFS_REQUIRE(kind == 0 || kind == 1, LAFS_ERR_ARG, "%s: kind=%d (0 CosFace, 1 ArcFace)", who, kind);
  LAFS_REQUIRE(!(kind == 1 && label_b != nullptr), LAFS_ERR_ARG, "%s: ArcFace takes hard labels only", who);
  LAFS_REQUIRE((((uintptr_t)e_hat | (uintptr_t)w_hat) & 15u) == 0, LAFS_ERR_ARG, "%s: operands must be 16-byte aligned", who);
  int rc = plan_head(B, C_local, D, hl);
  if (rc) return rc;
  plan_pair<MODE>(hl, C_local, D);
  rc = TmaEncoder::bf16_2d_sw128(te, e_hat, (uint64_t)B, (uint64_t)D, (uint64_t)D * 2, kBM);
  if (rc) return rc;
  rc = TmaEncoder::bf16_2d_sw128(tw, w_hat, (uint64_t)C_local, (uint64_t)D, (uint64_t)D * 2,
                                 (uint32_t)(hl->pair ? hl->BN / 2 : hl->BN));
  if (rc) return rc;
  *p = HeadParams{};
  p->B = B; p->C_local = C_local; p->D = D; p->class_lo = class_lo;
  p->kch = D / 64; p->stages = hl->stages; p->nranges = hl->nranges; p->nchunks = hl->nchunks;
  p->s = s; p->m = m; p->lam = lam; p->kind = kind;
  p->cos_m = cosf(m); p->sin_m = sinf(m);
  p->th = cosf(3.14159265358979323846f - m); p->mm = sinf(3.14159265358979323846f - m) * m;
  p->label_a = label_a; p->label_b = label_b;
  if (const char* dbg = getenv("LAFS_HEAD_DEBUG")) p->debug = atoi(dbg);
  return LAFS_OK;
}

}  // namespace lafs

using namespace lafs;

extern "C" int lafs_normalize_rows(const void* x, int dtype, int R, int D, void* out_bf16, float* inv_norm,
                                   lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(x)) return brc;
  LAFS_REQUIRE(x && out_bf16 && R >= 0 && D > 0, LAFS_ERR_ARG, "lafs_normalize_rows: bad argument");
  LAFS_REQUIRE(dtype >= 0 && dtype <= 2, LAFS_ERR_ARG, "lafs_normalize_rows: dtype=%d", dtype);
  if (R == 0) return LAFS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = (R + 7) / 8;
  if (dtype == LAFS_F32) launch_pdl((normalize_rows_kernel<float>), dim3(grid), dim3(256), (size_t)(0), st, (const float*)x, R, D, (__nv_bfloat16*)out_bf16, inv_norm);
  else if (dtype == LAFS_BF16) launch_pdl((normalize_rows_kernel<__nv_bfloat16>), dim3(grid), dim3(256), (size_t)(0), st, (const __nv_bfloat16*)x, R, D, (__nv_bfloat16*)out_bf16, inv_norm);
  else launch_pdl((normalize_rows_kernel<__half>), dim3(grid), dim3(256), (size_t)(0), st, (const __half*)x, R, D, (__nv_bfloat16*)out_bf16, inv_norm);
  return check_launch("lafs_normalize_rows");
}

extern "C" size_t lafs_head_workspace_bytes(int B, int C_local, int D) {
  HeadLaunch hl;
  if (B <= 0 || C_local <= 0 || plan_head(B, C_local, D, &hl) != LAFS_OK) return 0;
  // two partial records per (row, range): the shape-only plan bounds the pair plan's range count
  return (size_t)hl.mtiles * kBM * hl.nranges * 2 * 4 * sizeof(float);
}

extern "C" int lafs_head_fwd(const void* e_hat, const void* w_hat, const int64_t* label_a, const int64_t* label_b,
                             float lam, int B, int C_local, int D, int class_lo, float s, float m, int kind,
                             float* row_stats, void* workspace, size_t workspace_bytes, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(e_hat)) return brc;
  HeadParams p; HeadLaunch hl; CUtensorMap te, tw;
  int rc = head_common<HEAD_STATS>(e_hat, w_hat, label_a, label_b, lam, B, C_local, D, class_lo, s, m, kind, &p, &hl, &te, &tw, "lafs_head_fwd");
  if (rc) return rc;
  LAFS_REQUIRE(row_stats && workspace, LAFS_ERR_ARG, "lafs_head_fwd: null output");
  const size_t need = (size_t)hl.mtiles * kBM * hl.nranges * 2 * 4 * sizeof(float);
  LAFS_REQUIRE(workspace_bytes >= need, LAFS_ERR_WORKSPACE, "lafs_head_fwd: workspace %zu < %zu", workspace_bytes, need);
  p.part = (float*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  rc = launch_head_any<HEAD_STATS>(te, tw, p, hl, st);
  if (rc) return rc;
  launch_pdl((head_merge_kernel), dim3((B + 7) / 8), dim3(256), (size_t)(0), st, p.part, B, 2 * hl.nranges, 4, (long long)hl.nranges * 8, row_stats);
  return check_launch("lafs_head_fwd/merge");
}

extern "C" int lafs_head_logits(const void* e_hat, const void* w_hat, const int64_t* label_a, const int64_t* label_b,
                                float lam, int B, int C_local, int D, int class_lo, float s, float m, int kind,
                                float* logits, long long ldc, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(e_hat)) return brc;
  HeadParams p; HeadLaunch hl; CUtensorMap te, tw;
  int rc = head_common<HEAD_LOGITS>(e_hat, w_hat, label_a, label_b, lam, B, C_local, D, class_lo, s, m, kind, &p, &hl, &te, &tw, "lafs_head_logits");
  if (rc) return rc;
  LAFS_REQUIRE(logits && ldc >= C_local, LAFS_ERR_ARG, "lafs_head_logits: bad output");
  p.logits = logits; p.ldc = ldc;
  cudaStream_t st = (cudaStream_t)stream;
  return launch_head_any<HEAD_LOGITS>(te, tw, p, hl, st);
}

extern "C" int lafs_head_merge(const float* parts, int nparts, int B, float* row_stats, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(parts)) return brc;
  LAFS_REQUIRE(parts && row_stats && nparts > 0 && B > 0, LAFS_ERR_ARG, "lafs_head_merge: bad argument");
  launch_pdl((head_merge_kernel), dim3((B + 7) / 8), dim3(256), (size_t)(0), (cudaStream_t)stream, parts, B, nparts, (long long)B * 4, 4, row_stats);
  return check_launch("lafs_head_merge");
}

extern "C" int lafs_head_loss(const float* row_stats, const int64_t* label_a, const int64_t* label_b, float lam, int B,
                              float* row_lse2, float* loss_out, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(row_stats)) return brc;
  LAFS_REQUIRE(row_stats && label_a && row_lse2 && loss_out && B > 0, LAFS_ERR_ARG, "lafs_head_loss: bad argument");
  launch_pdl((head_loss_kernel), dim3(1), dim3(256), (size_t)(0), (cudaStream_t)stream, row_stats, label_a, label_b, lam, B, row_lse2, loss_out);
  return check_launch("lafs_head_loss");
}

extern "C" int lafs_head_grad_logits(const void* e_hat, const void* w_hat, const int64_t* label_a, const int64_t* label_b,
                                     float lam, int B, int C_local, int D, int class_lo, float s, float m, int kind,
                                     const float* row_lse2, const float* grad_out, float gscale, void* grad_bf16,
                                     long long ldg, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(e_hat)) return brc;
  HeadParams p; HeadLaunch hl; CUtensorMap te, tw;
  int rc = head_common<HEAD_GRAD>(e_hat, w_hat, label_a, label_b, lam, B, C_local, D, class_lo, s, m, kind, &p, &hl, &te, &tw, "lafs_head_grad_logits");
  if (rc) return rc;
  LAFS_REQUIRE(row_lse2 && grad_out && grad_bf16 && ldg >= C_local, LAFS_ERR_ARG, "lafs_head_grad_logits: bad argument");
  p.row_lse2 = row_lse2; p.grad = (__nv_bfloat16*)grad_bf16; p.ldg = ldg; p.gscale = gscale; p.grad_out = grad_out;
  p.g256 = (ldg % 16 == 0 && ((uintptr_t)grad_bf16 & 31u) == 0) ? 1 : 0;
  cudaStream_t st = (cudaStream_t)stream;
  return launch_head_any<HEAD_GRAD>(te, tw, p, hl, st);
}

/* lafs_head_grad_logits plus the per-class partial dots
 * tpart[(mtile*4 + quarter)*ldt + c] = sum over that 32-row group of G[b,c]*cos[b,c]; summing the
 * 4*ceil(B/128) partials of a class gives <w_hat_c, dW_hat_c>, which lafs_head_bwd_weight_t turns into the
 * F.normalize Jacobian on the tensor core.  ldt >= C_local rounded up to 32; tpart is fully overwritten for
 * classes < round32(C_local). */
extern "C" int lafs_head_grad_logits_t(const void* e_hat, const void* w_hat, const int64_t* label_a, const int64_t* label_b,
                                       float lam, int B, int C_local, int D, int class_lo, float s, float m, int kind,
                                       const float* row_lse2, const float* grad_out, float gscale, void* grad_bf16,
                                       long long ldg, float* tpart, long long ldt, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(e_hat)) return brc;
  HeadParams p; HeadLaunch hl; CUtensorMap te, tw;
  int rc = head_common<HEAD_GRAD_T>(e_hat, w_hat, label_a, label_b, lam, B, C_local, D, class_lo, s, m, kind, &p, &hl, &te, &tw, "lafs_head_grad_logits_t");
  if (rc) return rc;
  LAFS_REQUIRE(row_lse2 && grad_out && grad_bf16 && ldg >= C_local, LAFS_ERR_ARG, "lafs_head_grad_logits_t: bad argument");
  LAFS_REQUIRE(tpart && ldt >= ((C_local + 31) / 32) * 32 && ((uintptr_t)tpart & 127u) == 0 && ldt % 32 == 0, LAFS_ERR_ARG,
               "lafs_head_grad_logits_t: tpart must be 128-byte aligned with ldt a multiple of 32 >= round32(C_local)");
  p.row_lse2 = row_lse2; p.grad = (__nv_bfloat16*)grad_bf16; p.ldg = ldg; p.gscale = gscale; p.grad_out = grad_out;
  p.g256 = (ldg % 16 == 0 && ((uintptr_t)grad_bf16 & 31u) == 0) ? 1 : 0;
  p.part = tpart; p.ldc = ldt;
  return launch_head_any<HEAD_GRAD_T>(te, tw, p, hl, (cudaStream_t)stream);
}
