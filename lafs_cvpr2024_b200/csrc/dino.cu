// (2) DINO loss: teacher centring + temperature softmax, student log-softmax over all crops,
// cross-entropy and the centre column-sum in ONE streaming pass over the logits, plus the
// gradient pass and the centre EMA.  Replaces DINOLoss.forward/update_center
// (lafs_train.py:643-679), which launches ~60-100 kernels each materialising [B,K] fp32.
//
// Single-pass identity (sum_k q = 1):
//   loss = 1/(n_terms*B) * sum_b [ sum_v n_v*lse(s_v/ts) - (1/ts) * sum_iq sum_k q_iq,k*(S_k - s_iq,k) ]
//   q_iq = softmax((t_iq - c)/tt), S_k = sum_v s_v,k, n_v = #{iq != v}, n_terms = 2*ncrops-2.
//
// Work decomposition (HBM-bound, coalesced, warp-shuffle reduced):
//   * a lane owns NC = U*VEC fixed columns of a K-slice and keeps the centre and the
//     column-sum accumulators for them in registers;
//   * a warp walks over samples; per sample it loads the 2 teacher + ncrops student row
//     segments with 128-bit streaming loads (next sample prefetched), finds the per-row slice
//     maxima with one REDUX each, and produces per-(sample,slice) partial statistics;
//   * partials are merged across slices by dino_rows_finalize (deterministic, no atomics).
// All math in fp32, exponentials in the log2 domain (ex2.approx).
#include <stdlib.h>
#include "common.cuh"
#include "../../include/lafs_b200.h"

namespace lafs {

constexpr int kDinoThreads = 256;
constexpr int kDinoWarps = kDinoThreads / 32;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kMaxGroups = 32;

template <typename T> struct VecOf { static constexpr int VEC = 16 / sizeof(T); };

template <typename T>
__device__ __forceinline__ void unpack(const uint4& u, float* v) {
  if constexpr (sizeof(T) == 4) {
    v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y);
    v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
  } else {
    v[0] = Half2Ops<T>::lo(u.x); v[1] = Half2Ops<T>::hi(u.x);
    v[2] = Half2Ops<T>::lo(u.y); v[3] = Half2Ops<T>::hi(u.y);
    v[4] = Half2Ops<T>::lo(u.z); v[5] = Half2Ops<T>::hi(u.z);
    v[6] = Half2Ops<T>::lo(u.w); v[7] = Half2Ops<T>::hi(u.w);
  }
}
template <typename T>
__device__ __forceinline__ uint4 pack(const float* v) {
  uint4 u;
  if constexpr (sizeof(T) == 4) {
    u.x = __float_as_uint(v[0]); u.y = __float_as_uint(v[1]);
    u.z = __float_as_uint(v[2]); u.w = __float_as_uint(v[3]);
  } else {
    u.x = Half2Ops<T>::pack(v[0], v[1]); u.y = Half2Ops<T>::pack(v[2], v[3]);
    u.z = Half2Ops<T>::pack(v[4], v[5]); u.w = Half2Ops<T>::pack(v[6], v[7]);
  }
  return u;
}

// record of partial statistics per (sample, slice):
//   [0..2] teacher view 0: max (log2 domain), Z, A      [3..5] teacher view 1
//   [6+2v], [7+2v] student crop v: max, sum
__host__ __device__ constexpr int rec_floats(int ncrops) { return 6 + 2 * ncrops; }

template <typename T, int NCROPS, int U>
struct Rows {
  uint4 t[2][U];
  uint4 s[NCROPS][U];
};

constexpr int kChunk = 8;   // samples between two CTA-level pre-merges of the 8 warps' slice records

// Merge the 8 slice records of one sample (one per warp of the CTA, in shared memory) into ONE record of the
// same format and store it: the finishing kernel then reads 8x fewer records.  Whole warp; lanes l and
// l + 8k hold record l & 7, the xor butterflies stay inside 8-lane groups, so every lane ends with the result.
template <int NCROPS>
__device__ __forceinline__ void merge_cta_records(const float* __restrict__ recs /*[8][REC]*/, float* __restrict__ out,
                                                  int lane) {
  constexpr int REC = rec_floats(NCROPS);
  const float* r = recs + (lane & 7) * REC;
  float res[REC];
#pragma unroll
  for (int iq = 0; iq < 2; ++iq) {
    const float ml = r[3 * iq];
    float m = ml;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float f = ex2(ml - m);                  // empty record: ml = -inf -> 0
    float z = r[3 * iq + 1] * f, a = r[3 * iq + 2] * f;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      z += __shfl_xor_sync(0xffffffffu, z, o);
      a += __shfl_xor_sync(0xffffffffu, a, o);
    }
    res[3 * iq] = m; res[3 * iq + 1] = z; res[3 * iq + 2] = a;
  }
#pragma unroll
  for (int v = 0; v < NCROPS; ++v) {
    const float ml = r[6 + 2 * v];
    float m = ml;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float z = r[7 + 2 * v] * ex2(ml - m);
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
    res[6 + 2 * v] = m; res[7 + 2 * v] = z;
  }
  float val = 0.f;
#pragma unroll
  for (int j = 0; j < REC; ++j)
    if (lane == j) val = res[j];
  if (lane < REC) out[lane] = val;                // one coalesced store of the merged record
}

// RAGGED: the last K-slice is partially filled (K % (32*VEC) != 0); lanes past K are masked.
// The common case (K = 65536) compiles without any per-element predication.
template <typename T, int NCROPS, bool RAGGED>
__global__ void __launch_bounds__(kDinoThreads, 2)
dino_fwd_partial(const T* __restrict__ student, const T* __restrict__ teacher,
                 const float* __restrict__ center, int B, int K, float a_s, float a_t,
                 int nslices, int ngroups, float* __restrict__ part, float* __restrict__ colsum_part,
                 int b_begin, int b_count, unsigned* __restrict__ done_counter) {
  pdl_wait();
  constexpr int VEC = VecOf<T>::VEC;
  constexpr int NC = VEC;
  constexpr int REC = rec_floats(NCROPS);
  // slice records of the current / previous chunk of samples: [buffer][sample in chunk][warp][REC]
  __shared__ float srec[2][kChunk][kDinoWarps][REC];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // arrival counter of dino_finish (the next kernel on the stream): reset here so that the workspace
  // needs no initialisation by the caller
  if (done_counter != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *done_counter = 0u;
  // The 8 warps of a CTA own 8 ADJACENT slices and walk over the same samples at the same pace, so
  // a CTA touches 8 x 512 B = 4 KB contiguous bytes of every row it reads (DRAM page locality: with
  // one 512 B segment per row per CTA the same kernel ran at 40 % of the HBM rate).
  const int slice = blockIdx.x * kDinoWarps + warp, group = blockIdx.y;
  // a warp past the last slice (K not a multiple of 8 slices) computes nothing but keeps the CTA's barriers
  // balanced; its records are empty (max = -inf, sums = 0)
  const bool active = slice < nslices;
  const int nparts = (int)gridDim.x;              // merged records per sample = CTAs along K
  // this launch covers samples [b_begin, b_begin + b_count) (the whole batch, or one L2-sized wave)
  const int b_lo = b_begin + (int)(((long long)b_count * group) / ngroups);
  const int b_hi = b_begin + (int)(((long long)b_count * (group + 1)) / ngroups);

  const int col = slice * (32 * VEC) + lane * VEC;
  const bool ok = active && (!RAGGED || col < K);   // K % VEC == 0 is checked by the host
  const int colc = ok ? col : 0;                 // masked lanes read a valid address and ignore it
  float cen[NC], csum[NC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    cen[j] = __ldg(center + colc + j) * a_t;
    csum[j] = 0.f;
  }

  // Register-rotating software pipeline: `rows` always holds the not-yet-consumed segments of
  // the current sample; as soon as a row has been unpacked its registers are re-filled with the
  // same row of the warp's next sample, so ncrops+2 128-bit loads stay in flight per lane.
  // Row pointers are advanced by a constant byte stride instead of being recomputed.
  int b = b_lo;
  // address of (row r, this lane's columns) = base + off + r*row_bytes: one IMAD.WIDE per load,
  // `off` advanced by a constant stride per sample
  const unsigned row_bytes = (unsigned)((size_t)K * sizeof(T));
  const size_t step = (size_t)row_bytes;
  size_t off = (size_t)b * row_bytes + (size_t)colc * sizeof(T);
  const char* tbase = reinterpret_cast<const char*>(teacher);
  const char* sbase = reinterpret_cast<const char*>(student);
  auto t_ptr = [&](int iq) { return tbase + (off + (size_t)((unsigned)(iq * B)) * row_bytes); };
  auto s_ptr = [&](int v) { return sbase + (off + (size_t)((unsigned)(v * B)) * row_bytes); };
  uint4 rt[2], rs[NCROPS];
  if (active && b < b_hi) {
#pragma unroll
    for (int iq = 0; iq < 2; ++iq) rt[iq] = ld_stream_u4(t_ptr(iq));
#pragma unroll
    for (int v = 0; v < NCROPS; ++v) rs[v] = ld_stream_u4(s_ptr(v));
  }
  off += step;   // `off` now addresses the warp's NEXT sample
  for (; b < b_hi; ++b) {
    const int ci = (b - b_lo) % kChunk, cbuf = ((b - b_lo) / kChunk) & 1;
    float* dst = &srec[cbuf][ci][warp][0];        // this warp's record of sample b
   if (!active) {
    if (lane < REC) dst[lane] = ((lane < 6 ? lane % 3 == 0 : (lane & 1) == 0)) ? -INFINITY : 0.f;
   } else {
    const bool has_next = b + 1 < b_hi;
    // the register pipeline covers one iteration of latency; pull the sample after next into L2
    if (b + 2 < b_hi) {
#pragma unroll
      for (int iq = 0; iq < 2; ++iq) prefetch_l2(t_ptr(iq) + step);
#pragma unroll
      for (int v = 0; v < NCROPS; ++v) prefetch_l2(s_ptr(v) + step);
    }

    // partial record of this (sample, slice) in shared memory; lane 0 stores each entry as soon as it is final
    float zpart[2];
    float e[2][NC];  // teacher: first x (log2-domain logits), then exp2(x - max)
    // ---- teacher rows ------------------------------------------------------------------
#pragma unroll
    for (int iq = 0; iq < 2; ++iq) {
      float tv[VEC];
      unpack<T>(rt[iq], tv);
      if (has_next) rt[iq] = ld_stream_u4(t_ptr(iq));
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        if (!RAGGED || ok) csum[j] += tv[j];
        float x = fmaf(tv[j], a_t, -cen[j]);
        if (RAGGED && !ok) x = -INFINITY;
        e[iq][j] = x;
        mx = fmaxf(mx, x);
      }
      mx = warp_max_redux(mx);
      float z = 0.f;
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        e[iq][j] = ex2(e[iq][j] - mx);
        z += e[iq][j];
      }
      if (lane == 0) dst[3 * iq] = mx;
      zpart[iq] = z;
    }
    // ---- student rows ------------------------------------------------------------------
    float S[NC];
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int v = 0; v < NCROPS; ++v) {
      float sv[NC];
      unpack<T>(rs[v], sv);
      if (has_next) rs[v] = ld_stream_u4(s_ptr(v));
      float mx = sv[0];
#pragma unroll
      for (int j = 1; j < VEC; ++j) mx = fmaxf(mx, sv[j]);
      if (RAGGED && !ok) mx = -INFINITY;
      mx = warp_max_redux(mx) * a_s;  // a_s > 0: max commutes with the scaling
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        sum += ex2(fmaf(sv[j], a_s, -mx));
        S[j] = v == 0 ? sv[j] : S[j] + sv[j];
      }
      if (RAGGED && !ok) sum = 0.f;
      // the two self-view products are removed from the S-dot below
      if (v == 0) {
#pragma unroll
        for (int j = 0; j < NC; ++j) a0 = fmaf(e[0][j], -sv[j], a0);
      }
      if (v == 1) {
#pragma unroll
        for (int j = 0; j < NC; ++j) a1 = fmaf(e[1][j], -sv[j], a1);
      }
      sum = warp_sum_unit_terms(sum);   // terms are exp2(. - slice max) <= 1, at most 256 per slice
      if (lane == 0) *reinterpret_cast<float2*>(dst + 6 + 2 * v) = make_float2(mx, sum);
    }
#pragma unroll
    for (int j = 0; j < NC; ++j) {   // masked lanes: e == 0, so their (arbitrary) S does not count
      a0 = fmaf(e[0][j], S[j], a0);
      a1 = fmaf(e[1][j], S[j], a1);
    }
    // ---- warp reduction of the additive statistics ---------------------------------------
    zpart[0] = warp_sum_unit_terms(zpart[0]);
    zpart[1] = warp_sum_unit_terms(zpart[1]);
    // a0 and a1 share one butterfly: after the first exchange the lower half-warp carries a0 and the
    // upper half-warp a1 (5 shuffles instead of 10); lane 0 ends with sum(a0), lane 16 with sum(a1)
    {
      const bool up = (lane & 16) != 0;
      float mine = up ? a1 : a0;
      mine += __shfl_xor_sync(0xffffffffu, up ? a0 : a1, 16);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
      if (lane == 0) { dst[1] = zpart[0]; dst[2] = mine; dst[4] = zpart[1]; }
      if (lane == 16) dst[5] = mine;
    }
    off += step;
   }
    // ---- every kChunk samples (and after the last one): merge the 8 warps' records per sample ------------
    // One barrier per chunk; the buffers alternate, and a buffer is rewritten only after the NEXT chunk's
    // barrier, which every warp reaches after its merge of this chunk.
    if (ci == kChunk - 1 || b + 1 == b_hi) {
      __syncthreads();
      if (warp <= ci)
        merge_cta_records<NCROPS>(&srec[cbuf][warp][0][0],
                                  part + ((size_t)(b - ci + warp) * nparts + blockIdx.x) * REC, lane);
    }
  }

  // ---- column sums of this warp's samples (each warp owns its columns: no cross-warp reduction) ----
  if (ok) {
    float* dstc = colsum_part + (size_t)group * K + col;
#pragma unroll
    for (int j = 0; j < VEC; j += 4)
      *reinterpret_cast<float4*>(dstc + j) = make_float4(csum[j], csum[j + 1], csum[j + 2], csum[j + 3]);
  }
}

// One warp per sample: merge the slice partials, emit row statistics and the sample's loss.
constexpr int kFinalizeWarps = 2;   // small blocks: B/2 CTAs keep every SM busy on this latency-bound merge
// One warp per sample: merge the slice partials, emit row statistics and the sample's loss.
// Two passes over the (L1-resident) records: global maxima first, then sums rescaled to them --
// independent FMAs instead of a serial online-softmax chain.
template <int NCROPS>
__global__ void __launch_bounds__(kFinalizeWarps * 32)
dino_rows_finalize(const float* __restrict__ part, int B, int nslices, float inv_ts,
                   float* __restrict__ row_stats, float* __restrict__ sample_loss, int b_begin, int b_count) {
  pdl_wait();
  constexpr int REC = rec_floats(NCROPS);
  constexpr int NR = 2 + NCROPS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int bi = blockIdx.x * kFinalizeWarps + warp;
  if (bi >= b_count) return;
  const int b = b_begin + bi;
  const float* base = part + (size_t)b * nslices * REC;
  float m[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) m[i] = -INFINITY;
  for (int s = lane; s < nslices; s += 32) {
    const float* r = base + (size_t)s * REC;
    m[0] = fmaxf(m[0], r[0]);
    m[1] = fmaxf(m[1], r[3]);
#pragma unroll
    for (int v = 0; v < NCROPS; ++v) m[2 + v] = fmaxf(m[2 + v], r[6 + 2 * v]);
  }
#pragma unroll
  for (int i = 0; i < NR; ++i) m[i] = warp_max(m[i]);
  float z[NR], a[2];
#pragma unroll
  for (int i = 0; i < NR; ++i) z[i] = 0.f;
  a[0] = a[1] = 0.f;
  for (int s = lane; s < nslices; s += 32) {
    const float* r = base + (size_t)s * REC;
#pragma unroll
    for (int iq = 0; iq < 2; ++iq) {
      const float f = ex2(r[3 * iq] - m[iq]);
      z[iq] = fmaf(r[3 * iq + 1], f, z[iq]);
      a[iq] = fmaf(r[3 * iq + 2], f, a[iq]);
    }
#pragma unroll
    for (int v = 0; v < NCROPS; ++v) z[2 + v] = fmaf(r[7 + 2 * v], ex2(r[6 + 2 * v] - m[2 + v]), z[2 + v]);
  }
#pragma unroll
  for (int i = 0; i < NR; ++i) z[i] = warp_sum(z[i]);
  a[0] = warp_sum(a[0]);
  a[1] = warp_sum(a[1]);
  if (lane == 0) {
    float loss = 0.f;
#pragma unroll
    for (int v = 0; v < NCROPS; ++v) {
      const float l2 = m[2 + v] + lg2(z[2 + v]);  // log2-domain lse of s_v/ts
      row_stats[(size_t)v * B + b] = l2;
      loss += (v < 2 ? 1.f : 2.f) * l2;
    }
    loss *= kLn2;
#pragma unroll
    for (int iq = 0; iq < 2; ++iq) {
      row_stats[(size_t)(NCROPS + iq) * B + b] = m[iq] + lg2(z[iq]);
      loss -= inv_ts * (a[iq] / z[iq]);
    }
    sample_loss[b] = loss;
  }
}

// block 0: fixed-order sum of the per-sample losses; all blocks: merge column-sum partials.
__global__ void __launch_bounds__(kDinoThreads)
dino_tail(const float* __restrict__ sample_loss, int B, float inv_norm, float* __restrict__ loss_out,
          const float* __restrict__ colsum_part, int ngroups, int K, float* __restrict__ colsum_out,
          const float* __restrict__ center, float* __restrict__ center_out, float count, float mom, float om) {
  pdl_wait();
  const int k = blockIdx.x * kDinoThreads + threadIdx.x;
  if (k < K) {
    float acc = 0.f;
    for (int g = 0; g < ngroups; ++g) acc += colsum_part[(size_t)g * K + k];
    colsum_out[k] = acc;
    // single-process case: the centre EMA (lafs_train.py:676-679) rides along (same arithmetic as
    // center_ema_kernel); with several ranks the caller all-reduces colsum first
    if (center_out != nullptr)
      center_out[k] = __fadd_rn(__fmul_rn(center[k], mom), __fmul_rn(__fdiv_rn(acc, count), om));
  }
  if (blockIdx.x == 0) {
    __shared__ float red[kDinoThreads];
    float acc = 0.f;
    for (int i = threadIdx.x; i < B; i += kDinoThreads) acc += sample_loss[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = kDinoThreads / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) *loss_out = red[0] * inv_norm;
  }
}

// Everything that follows the streaming pass, in ONE launch of 64-thread blocks:
//   blocks [0, nfin)      : one warp per sample -- the sample's slice records (contiguous, 72 B x
//                           nslices) are staged into shared memory with coalesced 128-bit loads, then
//                           merged in two passes exactly like dino_rows_finalize; the warp that
//                           arrives last (device-wide counter) adds the per-sample losses in a fixed
//                           order, so the result does not depend on the arrival order;
//   blocks [nfin, gridDim): column-sum partials -> column sums (+ the centre EMA when one process
//                           owns the whole batch), 4 columns per thread.
template <int NCROPS>
__global__ void __launch_bounds__(64)
dino_finish(const float* __restrict__ part, int B, int nslices, float inv_ts, float* __restrict__ row_stats,
            float* __restrict__ sample_loss, int nfin, float inv_norm, float* __restrict__ loss_out,
            unsigned* __restrict__ done_counter, const float* __restrict__ colsum_part, int ngroups, int K,
            float* __restrict__ colsum_out, const float* __restrict__ center, float* __restrict__ center_out,
            float count, float mom, float om) {
  pdl_wait();
  constexpr int REC = rec_floats(NCROPS);
  constexpr int NR = 2 + NCROPS;
  extern __shared__ __align__(16) float fin_smem[];
  if ((int)blockIdx.x >= nfin) {
    const int k = (((int)blockIdx.x - nfin) * 64 + (int)threadIdx.x) * 4;
    if (k < K) {                                        // K % 4 == 0 (checked by the host)
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int g = 0; g < ngroups; ++g) {
        const float4 v = *reinterpret_cast<const float4*>(colsum_part + (size_t)g * K + k);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      *reinterpret_cast<float4*>(colsum_out + k) = acc;
      if (center_out != nullptr) {                      // same arithmetic as center_ema_kernel
        const float4 c = *reinterpret_cast<const float4*>(center + k);
        float4 o;
        o.x = __fadd_rn(__fmul_rn(c.x, mom), __fmul_rn(__fdiv_rn(acc.x, count), om));
        o.y = __fadd_rn(__fmul_rn(c.y, mom), __fmul_rn(__fdiv_rn(acc.y, count), om));
        o.z = __fadd_rn(__fmul_rn(c.z, mom), __fmul_rn(__fdiv_rn(acc.z, count), om));
        o.w = __fadd_rn(__fmul_rn(c.w, mom), __fmul_rn(__fdiv_rn(acc.w, count), om));
        *reinterpret_cast<float4*>(center_out + k) = o;
      }
    }
    return;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int b = blockIdx.x * 2 + warp;
  if (b < B) {
    const int nrec = nslices * REC;
    float* rec = fin_smem + (size_t)warp * ((nrec + 3) & ~3);
    const float* src = part + (size_t)b * nrec;
    if ((nrec & 3) == 0) {
      for (int i = lane; i < nrec / 4; i += 32)
        *reinterpret_cast<float4*>(rec + 4 * i) = *reinterpret_cast<const float4*>(src + 4 * i);
    } else {
      for (int i = lane; i < nrec; i += 32) rec[i] = src[i];
    }
    __syncwarp();
    float m[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) m[i] = -INFINITY;
    for (int s = lane; s < nslices; s += 32) {
      const float* r = rec + s * REC;
      m[0] = fmaxf(m[0], r[0]);
      m[1] = fmaxf(m[1], r[3]);
#pragma unroll
      for (int v = 0; v < NCROPS; ++v) m[2 + v] = fmaxf(m[2 + v], r[6 + 2 * v]);
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) m[i] = warp_max(m[i]);
    float z[NR], a[2];
#pragma unroll
    for (int i = 0; i < NR; ++i) z[i] = 0.f;
    a[0] = a[1] = 0.f;
    for (int s = lane; s < nslices; s += 32) {
      const float* r = rec + s * REC;
#pragma unroll
      for (int iq = 0; iq < 2; ++iq) {
        const float f = ex2(r[3 * iq] - m[iq]);
        z[iq] = fmaf(r[3 * iq + 1], f, z[iq]);
        a[iq] = fmaf(r[3 * iq + 2], f, a[iq]);
      }
#pragma unroll
      for (int v = 0; v < NCROPS; ++v) z[2 + v] = fmaf(r[7 + 2 * v], ex2(r[6 + 2 * v] - m[2 + v]), z[2 + v]);
    }
#pragma unroll
    for (int i = 0; i < NR; ++i) z[i] = warp_sum(z[i]);
    a[0] = warp_sum(a[0]);
    a[1] = warp_sum(a[1]);
    if (lane == 0) {
      float loss = 0.f;
#pragma unroll
      for (int v = 0; v < NCROPS; ++v) {
        const float l2 = m[2 + v] + lg2(z[2 + v]);  // log2-domain lse of s_v/ts
        row_stats[(size_t)v * B + b] = l2;
        loss += (v < 2 ? 1.f : 2.f) * l2;
      }
      loss *= kLn2;
#pragma unroll
      for (int iq = 0; iq < 2; ++iq) {
        row_stats[(size_t)(NCROPS + iq) * B + b] = m[iq] + lg2(z[iq]);
        loss -= inv_ts * (a[iq] / z[iq]);
      }
      sample_loss[b] = loss;
    }
  }
  // ---- the warp that arrives last sums the per-sample losses (fixed order) ---------------------------
  int last = 0;
  if (lane == 0) {
    __threadfence();
    last = atomicAdd(done_counter, 1u) == (unsigned)(2 * nfin - 1);
  }
  last = __shfl_sync(0xffffffffu, last, 0);
  if (last) {
    __threadfence();
    float acc = 0.f;
    for (int i = lane; i < B; i += 32) acc += __ldcg(sample_loss + i);
    acc = warp_sum(acc);
    if (lane == 0) *loss_out = acc * inv_norm;
  }
}

// Gradient pass: ds[v,b,k] = g * ( n_v * softmax(s_v/ts)_k - sum_{iq != v} q_iq,k ),
// g = grad_out / (ts * n_terms * B).  Purely element-wise given the saved row statistics.
template <typename T, int NCROPS>
__global__ void __launch_bounds__(kDinoThreads)
dino_bwd_kernel(const T* __restrict__ student, const T* __restrict__ teacher,
                const float* __restrict__ center, const float* __restrict__ row_stats,
                const float* __restrict__ grad_out, int B, int K, float a_s, float a_t,
                float gcoef, T* __restrict__ grad_student, int b_begin) {
  pdl_wait();
  constexpr int VEC = VecOf<T>::VEC;
  const int b = b_begin + blockIdx.y;
  const int col = (blockIdx.x * kDinoThreads + threadIdx.x) * VEC;
  if (col >= K) return;
  uint4 tv[2], sv[NCROPS];
#pragma unroll
  for (int iq = 0; iq < 2; ++iq) tv[iq] = ld_stream_u4(teacher + (size_t)(iq * B + b) * K + col);
#pragma unroll
  for (int v = 0; v < NCROPS; ++v) sv[v] = ld_stream_u4(student + (size_t)(v * B + b) * K + col);
  const float g = __ldg(grad_out) * gcoef;
  float q[2][VEC];
#pragma unroll
  for (int iq = 0; iq < 2; ++iq) {
    const float r = __ldg(row_stats + (size_t)(NCROPS + iq) * B + b);
    float t[VEC];
    unpack<T>(tv[iq], t);
#pragma unroll
    for (int j = 0; j < VEC; ++j)
      q[iq][j] = ex2(fmaf(t[j], a_t, -fmaf(__ldg(center + col + j), a_t, r)));
  }
#pragma unroll
  for (int v = 0; v < NCROPS; ++v) {
    const float l2 = __ldg(row_stats + (size_t)v * B + b);
    float s[VEC], d[VEC];
    unpack<T>(sv[v], s);
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      const float p = ex2(fmaf(s[j], a_s, -l2));
      float r;
      if (v == 0) r = p - q[1][j];
      else if (v == 1) r = p - q[0][j];
      else r = 2.f * p - q[0][j] - q[1][j];
      d[j] = g * r;
    }
    st_stream_u4(grad_student + (size_t)(v * B + b) * K + col, pack<T>(d));
  }
}

// Stand-alone teacher column sum for DINOLoss.update_center called outside forward
// (lafs_train.py:674).  Off the hot path: forward already produces the column sums.
template <typename T>
__global__ void __launch_bounds__(kDinoThreads)
colsum_partial_kernel(const T* __restrict__ x, int rows, int K, int ngroups, float* __restrict__ part) {
  pdl_wait();
  constexpr int VEC = VecOf<T>::VEC;
  const int col = (blockIdx.x * kDinoThreads + threadIdx.x) * VEC;
  if (col >= K) return;
  const int g = blockIdx.y;
  const int r_lo = (int)(((long long)rows * g) / ngroups), r_hi = (int)(((long long)rows * (g + 1)) / ngroups);
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
  for (int r = r_lo; r < r_hi; ++r) {
    float v[VEC];
    unpack<T>(ld_stream_u4(x + (size_t)r * K + col), v);
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] += v[j];
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) part[(size_t)g * K + col + j] = acc[j];
}
__global__ void colsum_combine_kernel(const float* __restrict__ part, int ngroups, int K, float* __restrict__ out) {
  pdl_wait();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float acc = 0.f;
  for (int g = 0; g < ngroups; ++g) acc += part[(size_t)g * K + k];
  out[k] = acc;
}

__global__ void center_ema_kernel(const float* __restrict__ center, const float* __restrict__ colsum,
                                  float count, float mom, float om, int K, float* __restrict__ center_out) {
  pdl_wait();
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  // reference: center*momentum + (batch_center/n)*(1-momentum): separately rounded fp32 ops,
  // IEEE division as on the reference's CPU path
  if (k < K) center_out[k] = __fadd_rn(__fmul_rn(center[k], mom), __fmul_rn(__fdiv_rn(colsum[k], count), om));
}

// ---- host side ---------------------------------------------------------------------------
struct DinoPlan {
  int U, cols_per_slice, nslices, ngroups;
  size_t off_part, off_colsum, off_sample, off_counter, total;
};

static DinoPlan make_plan(int B, int K, int ncrops, int elem_bytes) {
  DinoPlan p;
  const int vec = 16 / elem_bytes;
  p.U = 1;
  p.cols_per_slice = 32 * vec * p.U;
  p.nslices = (K + p.cols_per_slice - 1) / p.cols_per_slice;
  // choose the number of sample groups so that the grid fills whole waves (2 CTAs per SM)
  const int cap = kNumSMs * 2;
  int best = 1;
  double best_eff = 0.0;
  const int gmax = B < kMaxGroups ? B : kMaxGroups;
  const int xctas = (p.nslices + kDinoWarps - 1) / kDinoWarps;
  for (int g = 1; g <= gmax; ++g) {
    const long long n = (long long)xctas * g;
    const long long waves = (n + cap - 1) / cap;
    const double eff = (double)n / (double)(waves * cap);
    if (eff > best_eff + 1e-9 || (eff > best_eff - 0.02 && g > best)) {
      if (eff > best_eff) best_eff = eff;
      best = g;
    }
  }
  p.ngroups = best;
  if (const char* e = getenv("LAFS_DINO_GROUPS")) {   // development override
    const int g = atoi(e);
    if (g >= 1 && g <= kMaxGroups && g <= B) p.ngroups = g;
  }
  size_t o = 0;
  p.off_part = o;   o += (size_t)B * p.nslices * rec_floats(ncrops) * sizeof(float);
  o = (o + 255) & ~(size_t)255;
  p.off_colsum = o; o += (size_t)p.ngroups * K * sizeof(float);
  o = (o + 255) & ~(size_t)255;
  p.off_sample = o; o += (size_t)B * sizeof(float);
  o = (o + 255) & ~(size_t)255;
  p.off_counter = o; o += 256;
  p.total = (o + 255) & ~(size_t)255;
  return p;
}

// dino_finish: shared-memory staging of two samples' slice records; falls back to the two small
// kernels when that does not fit (huge out_dim with fp32 logits and many crops)
template <int NCROPS>
static bool launch_finish(const float* part, int B, int nslices, float inv_ts, float* row_stats, float* sample_loss,
                          float* loss_out, unsigned* counter, const float* colsum_part, int ngroups, int K,
                          float* colsum_out, const float* center, float* center_out, float mom, float om,
                          cudaStream_t st) {
  const size_t nrec = ((size_t)nslices * rec_floats(NCROPS) + 3) & ~(size_t)3;
  const size_t smem = 2 * nrec * sizeof(float);
  if (smem > 200 * 1024) return false;
  auto kern = dino_finish<NCROPS>;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  const int nfin = (B + 1) / 2;
  const int ncol = (K / 4 + 63) / 64;
  const float inv_norm = 1.f / ((float)(2 * NCROPS - 2) * (float)B);
  const float count = (float)(2 * B);
  if (launch_pdl(kern, dim3(nfin + ncol), dim3(64), smem, st, part, B, nslices, inv_ts, row_stats, sample_loss, nfin, inv_norm,
                 loss_out, counter, colsum_part, ngroups, K, colsum_out, center, center_out, count, mom, om) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return true;
}

template <typename T, int NCROPS>
static int launch_fwd(const void* student, const void* teacher, const float* center, int B, int K,
                      float inv_ts, float inv_tt, float* loss_out, float* row_stats,
                      float* colsum_out, char* ws, const DinoPlan& p, cudaStream_t st,
                      float* center_out, float mom, float om) {
  float* part = reinterpret_cast<float*>(ws + p.off_part);
  float* colsum_part = reinterpret_cast<float*>(ws + p.off_colsum);
  float* sample_loss = reinterpret_cast<float*>(ws + p.off_sample);
  unsigned* counter = reinterpret_cast<unsigned*>(ws + p.off_counter);
  dim3 grid((p.nslices + kDinoWarps - 1) / kDinoWarps, p.ngroups);
  if (K % p.cols_per_slice == 0)
    launch_pdl((dino_fwd_partial<T, NCROPS, false>), dim3(grid), dim3(kDinoThreads), (size_t)(0), st, 
        (const T*)student, (const T*)teacher, center, B, K, inv_ts * kLog2e, inv_tt * kLog2e,
        p.nslices, p.ngroups, part, colsum_part, 0, B, counter);
  else
    launch_pdl((dino_fwd_partial<T, NCROPS, true>), dim3(grid), dim3(kDinoThreads), (size_t)(0), st, 
        (const T*)student, (const T*)teacher, center, B, K, inv_ts * kLog2e, inv_tt * kLog2e,
        p.nslices, p.ngroups, part, colsum_part, 0, B, counter);
  const int nparts = (int)grid.x;   // the streaming kernel leaves ONE merged record per (sample, CTA along K)
  if (launch_finish<NCROPS>(part, B, nparts, inv_ts, row_stats, sample_loss, loss_out, counter, colsum_part,
                            p.ngroups, K, colsum_out, center, center_out, mom, om, st))
    return check_launch("lafs_dino_fwd");
  launch_pdl((dino_rows_finalize<NCROPS>), dim3((B + kFinalizeWarps - 1) / kFinalizeWarps), dim3(kFinalizeWarps * 32), (size_t)(0), st, 
      part, B, nparts, inv_ts, row_stats, sample_loss, 0, B);
  const float inv_norm = 1.f / ((float)(2 * NCROPS - 2) * (float)B);
  launch_pdl((dino_tail), dim3((K + kDinoThreads - 1) / kDinoThreads), dim3(kDinoThreads), (size_t)(0), st, 
      sample_loss, B, inv_norm, loss_out, colsum_part, p.ngroups, K, colsum_out, center, center_out,
      (float)(2 * B), mom, om);
  return check_launch("lafs_dino_fwd");
}

template <typename T, int NCROPS>
static int launch_bwd(const void* student, const void* teacher, const float* center,
                      const float* row_stats, const float* grad_out, int B, int K, float inv_ts,
                      float inv_tt, void* grad_student, cudaStream_t st) {
  constexpr int VEC = VecOf<T>::VEC;
  dim3 grid((K / VEC + kDinoThreads - 1) / kDinoThreads, B);
  const float gcoef = inv_ts / ((float)(2 * NCROPS - 2) * (float)B);
  launch_pdl((dino_bwd_kernel<T, NCROPS>), dim3(grid), dim3(kDinoThreads), (size_t)(0), st, 
      (const T*)student, (const T*)teacher, center, row_stats, grad_out, B, K, inv_ts * kLog2e,
      inv_tt * kLog2e, gcoef, (T*)grad_student, 0);
  return check_launch("lafs_dino_bwd");
}

// Forward AND backward from one entry point, optionally in waves of samples so that the gradient
// pass of a wave re-reads from L2 what its forward pass has just streamed.  Measured on B200
// (tools/dino_fused_sweep.py, B=256, K=65536): one wave 174 us, 128-sample waves 194 us, 64: 238 us,
// 32: 307 us -- the shorter per-warp pipelines and partially filled grids of small waves cost more
// than the L2 hits save, so the default is ONE wave (LAFS_DINO_WAVE overrides).  Used when the
// upstream gradient of the loss is known when the loss is computed (grad_out device scalar).
constexpr int kMaxColsumRows = 64;
struct FusedPlan {
  int wave, nwaves, ngroups, nslices, cols_per_slice;
  size_t off_part, off_colsum, off_sample, off_counter, total;
};
static FusedPlan make_fused_plan(int B, int K, int ncrops, int elem_bytes) {
  FusedPlan p;
  const int vec = 16 / elem_bytes;
  p.cols_per_slice = 32 * vec;
  p.nslices = (K + p.cols_per_slice - 1) / p.cols_per_slice;
  const double bytes_per_sample = (double)(ncrops + 2) * K * elem_bytes;
  int wave = B;
  (void)bytes_per_sample;
  if (const char* e = getenv("LAFS_DINO_WAVE")) wave = atoi(e);
  if (wave < 8) wave = 8;
  if (wave > B) wave = B;
  p.nwaves = (B + wave - 1) / wave;
  if (p.nwaves > 16) { p.nwaves = 16; }
  p.wave = (B + p.nwaves - 1) / p.nwaves;
  p.nwaves = (B + p.wave - 1) / p.wave;
  int g = kMaxColsumRows / p.nwaves;
  if (g > kMaxGroups) g = kMaxGroups;
  const int xctas = (p.nslices + kDinoWarps - 1) / kDinoWarps;
  const int want = (2 * kNumSMs + xctas - 1) / xctas;            // >= one full wave of CTAs
  if (g > want) g = want;
  if (g > p.wave) g = p.wave;
  if (const char* e = getenv("LAFS_DINO_GROUPS")) { const int v = atoi(e); if (v >= 1 && v <= g) g = v; }
  if (g < 1) g = 1;
  if (p.nwaves == 1) g = make_plan(B, K, ncrops, elem_bytes).ngroups;   // same grid as lafs_dino_fwd
  p.ngroups = g;
  size_t o = 0;
  p.off_part = o;   o += (size_t)B * p.nslices * rec_floats(ncrops) * sizeof(float);
  o = (o + 255) & ~(size_t)255;
  p.off_colsum = o; o += (size_t)p.nwaves * p.ngroups * K * sizeof(float);
  o = (o + 255) & ~(size_t)255;
  p.off_sample = o; o += (size_t)B * sizeof(float);
  o = (o + 255) & ~(size_t)255;
  p.off_counter = o; o += 256;
  p.total = (o + 255) & ~(size_t)255;
  return p;
}

template <typename T, int NCROPS>
static int launch_fused(const void* student, const void* teacher, const float* center, const float* grad_out,
                        int B, int K, float inv_ts, float inv_tt, float* loss_out, float* row_stats,
                        float* colsum_out, void* grad_student, char* ws, const FusedPlan& p, cudaStream_t st,
                        float* center_out, float mom, float om) {
  constexpr int VEC = VecOf<T>::VEC;
  float* part = reinterpret_cast<float*>(ws + p.off_part);
  float* colsum_part = reinterpret_cast<float*>(ws + p.off_colsum);
  float* sample_loss = reinterpret_cast<float*>(ws + p.off_sample);
  const float gcoef = inv_ts / ((float)(2 * NCROPS - 2) * (float)B);
  const bool ragged = K % p.cols_per_slice != 0;
  unsigned* counter = reinterpret_cast<unsigned*>(ws + p.off_counter);
  for (int w = 0; w < p.nwaves; ++w) {
    const int b0 = w * p.wave;
    const int bc = (B - b0) < p.wave ? (B - b0) : p.wave;
    dim3 grid((p.nslices + kDinoWarps - 1) / kDinoWarps, p.ngroups);
    float* cpart = colsum_part + (size_t)w * p.ngroups * K;
    if (!ragged)
      launch_pdl((dino_fwd_partial<T, NCROPS, false>), dim3(grid), dim3(kDinoThreads), (size_t)(0), st, 
          (const T*)student, (const T*)teacher, center, B, K, inv_ts * kLog2e, inv_tt * kLog2e,
          p.nslices, p.ngroups, part, cpart, b0, bc, counter);
    else
      launch_pdl((dino_fwd_partial<T, NCROPS, true>), dim3(grid), dim3(kDinoThreads), (size_t)(0), st, 
          (const T*)student, (const T*)teacher, center, B, K, inv_ts * kLog2e, inv_tt * kLog2e,
          p.nslices, p.ngroups, part, cpart, b0, bc, counter);
    const int nparts = (int)grid.x;
    if (p.nwaves == 1 &&
        launch_finish<NCROPS>(part, B, nparts, inv_ts, row_stats, sample_loss, loss_out, counter, colsum_part,
                              p.ngroups, K, colsum_out, center, center_out, mom, om, st)) {
      dim3 gb1((K / VEC + kDinoThreads - 1) / kDinoThreads, B);
      launch_pdl((dino_bwd_kernel<T, NCROPS>), dim3(gb1), dim3(kDinoThreads), (size_t)(0), st, 
          (const T*)student, (const T*)teacher, center, row_stats, grad_out, B, K, inv_ts * kLog2e,
          inv_tt * kLog2e, gcoef, (T*)grad_student, 0);
      return check_launch("lafs_dino_fwd_bwd");
    }
    launch_pdl((dino_rows_finalize<NCROPS>), dim3((bc + kFinalizeWarps - 1) / kFinalizeWarps), dim3(kFinalizeWarps * 32), (size_t)(0), st, 
        part, B, nparts, inv_ts, row_stats, sample_loss, b0, bc);
    dim3 gb((K / VEC + kDinoThreads - 1) / kDinoThreads, bc);
    launch_pdl((dino_bwd_kernel<T, NCROPS>), dim3(gb), dim3(kDinoThreads), (size_t)(0), st, 
        (const T*)student, (const T*)teacher, center, row_stats, grad_out, B, K, inv_ts * kLog2e,
        inv_tt * kLog2e, gcoef, (T*)grad_student, b0);
  }
  const float inv_norm = 1.f / ((float)(2 * NCROPS - 2) * (float)B);
  launch_pdl((dino_tail), dim3((K + kDinoThreads - 1) / kDinoThreads), dim3(kDinoThreads), (size_t)(0), st, 
      sample_loss, B, inv_norm, loss_out, colsum_part, p.nwaves * p.ngroups, K, colsum_out, center, center_out,
      (float)(2 * B), mom, om);
  return check_launch("lafs_dino_fwd_bwd");
}

#define LAFS_DINO_CROPS(M) M(2) M(3) M(4) M(5) M(6) M(7) M(8) M(9) M(10) M(11) M(12)

}  // namespace lafs

extern "C" size_t lafs_dino_workspace_bytes(int B, int K, int ncrops) {
  if (B <= 0 || K <= 0 || ncrops < 2 || ncrops > 12) return 0;
  // sized for the worst case over dtypes (fp32 has the most slices)
  return lafs::make_plan(B, K, ncrops, 4).total;
}

static int dino_check(const void* s, const void* t, const float* c, int B, int K, int ncrops, int dtype,
                      const char* who) {
  using namespace lafs;
  LAFS_REQUIRE(s && t && c, LAFS_ERR_ARG, "%s: null pointer", who);
  LAFS_REQUIRE(B > 0 && K > 0, LAFS_ERR_ARG, "%s: B=%d K=%d must be positive", who, B, K);
  LAFS_REQUIRE(ncrops >= 2 && ncrops <= 12, LAFS_ERR_ARG, "%s: ncrops=%d outside [2,12]", who, ncrops);
  LAFS_REQUIRE(dtype >= 0 && dtype <= 2, LAFS_ERR_ARG, "%s: dtype=%d", who, dtype);
  const int vec = dtype == LAFS_F32 ? 4 : 8;
  LAFS_REQUIRE(K % vec == 0, LAFS_ERR_ARG, "%s: K=%d must be a multiple of %d for this dtype", who, K, vec);
  LAFS_REQUIRE((((uintptr_t)s | (uintptr_t)t | (uintptr_t)c) & 15u) == 0, LAFS_ERR_ARG,
               "%s: pointers must be 16-byte aligned", who);
  return LAFS_OK;
}

extern "C" int lafs_dino_fwd(const void* student, const void* teacher, const float* center, int B, int K,
                             int ncrops, float inv_student_temp, float inv_teacher_temp, int dtype,
                             float* loss_out, float* row_stats, float* colsum_out, void* workspace,
                             size_t workspace_bytes, float* center_out, float momentum,
                             float one_minus_momentum, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(student)) return brc;
  using namespace lafs;
  int rc = dino_check(student, teacher, center, B, K, ncrops, dtype, "lafs_dino_fwd");
  if (rc) return rc;
  LAFS_REQUIRE(loss_out && row_stats && colsum_out && workspace, LAFS_ERR_ARG, "lafs_dino_fwd: null output");
  const DinoPlan p = make_plan(B, K, ncrops, dtype == LAFS_F32 ? 4 : 2);
  LAFS_REQUIRE(workspace_bytes >= p.total, LAFS_ERR_WORKSPACE, "lafs_dino_fwd: workspace %zu < %zu",
               workspace_bytes, p.total);
  LAFS_REQUIRE(((uintptr_t)workspace & 255u) == 0, LAFS_ERR_ARG, "lafs_dino_fwd: workspace must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
#define LAFS_CASE(N)                                                                                     \
  case N:                                                                                                \
    if (dtype == LAFS_F32)                                                                               \
      return launch_fwd<float, N>(student, teacher, center, B, K, inv_student_temp, inv_teacher_temp,    \
                                  loss_out, row_stats, colsum_out, ws, p, st, center_out, momentum,       \
                                  one_minus_momentum);                                                   \
    if (dtype == LAFS_BF16)                                                                              \
      return launch_fwd<__nv_bfloat16, N>(student, teacher, center, B, K, inv_student_temp,              \
                                          inv_teacher_temp, loss_out, row_stats, colsum_out, ws, p, st,  \
                                          center_out, momentum, one_minus_momentum);                     \
    return launch_fwd<__half, N>(student, teacher, center, B, K, inv_student_temp, inv_teacher_temp,     \
                                 loss_out, row_stats, colsum_out, ws, p, st, center_out, momentum,       \
                                 one_minus_momentum);
  switch (ncrops) { LAFS_DINO_CROPS(LAFS_CASE) }
#undef LAFS_CASE
  return LAFS_ERR_ARG;
}

extern "C" size_t lafs_dino_fused_workspace_bytes(int B, int K, int ncrops) {
  if (B <= 0 || K <= 0 || ncrops < 2 || ncrops > 12) return 0;
  return lafs::make_fused_plan(B, K, ncrops, 4).total + (size_t)lafs::kMaxColsumRows * K * sizeof(float);
}

extern "C" int lafs_dino_fwd_bwd(const void* student, const void* teacher, const float* center,
                                 const float* grad_out, int B, int K, int ncrops, float inv_student_temp,
                                 float inv_teacher_temp, int dtype, float* loss_out, float* row_stats,
                                 float* colsum_out, void* grad_student, void* workspace, size_t workspace_bytes,
                                 float* center_out, float momentum, float one_minus_momentum,
                                 lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(student)) return brc;
  using namespace lafs;
  int rc = dino_check(student, teacher, center, B, K, ncrops, dtype, "lafs_dino_fwd_bwd");
  if (rc) return rc;
  LAFS_REQUIRE(grad_out && loss_out && row_stats && colsum_out && grad_student && workspace, LAFS_ERR_ARG,
               "lafs_dino_fwd_bwd: null pointer");
  LAFS_REQUIRE(((uintptr_t)grad_student & 15u) == 0 && ((uintptr_t)workspace & 255u) == 0, LAFS_ERR_ARG,
               "lafs_dino_fwd_bwd: misaligned grad_student / workspace");
  const FusedPlan p = make_fused_plan(B, K, ncrops, dtype == LAFS_F32 ? 4 : 2);
  LAFS_REQUIRE(workspace_bytes >= p.total, LAFS_ERR_WORKSPACE, "lafs_dino_fwd_bwd: workspace %zu < %zu",
               workspace_bytes, p.total);
  cudaStream_t st = (cudaStream_t)stream;
  char* ws = (char*)workspace;
#define LAFS_CASE(N)                                                                                          \
  case N:                                                                                                     \
    if (dtype == LAFS_F32)                                                                                    \
      return launch_fused<float, N>(student, teacher, center, grad_out, B, K, inv_student_temp,              \
                                    inv_teacher_temp, loss_out, row_stats, colsum_out, grad_student, ws, p,  \
                                    st, center_out, momentum, one_minus_momentum);                            \
    if (dtype == LAFS_BF16)                                                                                   \
      return launch_fused<__nv_bfloat16, N>(student, teacher, center, grad_out, B, K, inv_student_temp,      \
                                            inv_teacher_temp, loss_out, row_stats, colsum_out, grad_student, \
                                            ws, p, st, center_out, momentum, one_minus_momentum);             \
    return launch_fused<__half, N>(student, teacher, center, grad_out, B, K, inv_student_temp,               \
                                   inv_teacher_temp, loss_out, row_stats, colsum_out, grad_student, ws, p,   \
                                   st, center_out, momentum, one_minus_momentum);
  switch (ncrops) { LAFS_DINO_CROPS(LAFS_CASE) }
#undef LAFS_CASE
  return LAFS_ERR_ARG;
}

extern "C" int lafs_dino_bwd(const void* student, const void* teacher, const float* center,
                             const float* row_stats, const float* grad_out, int B, int K, int ncrops,
                             float inv_student_temp, float inv_teacher_temp, int dtype, void* grad_student,
                             lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(student)) return brc;
  using namespace lafs;
  int rc = dino_check(student, teacher, center, B, K, ncrops, dtype, "lafs_dino_bwd");
  if (rc) return rc;
  LAFS_REQUIRE(row_stats && grad_out && grad_student, LAFS_ERR_ARG, "lafs_dino_bwd: null pointer");
  LAFS_REQUIRE(((uintptr_t)grad_student & 15u) == 0, LAFS_ERR_ARG, "lafs_dino_bwd: grad_student misaligned");
  cudaStream_t st = (cudaStream_t)stream;
#define LAFS_CASE(N)                                                                                       \
  case N:                                                                                                  \
    if (dtype == LAFS_F32)                                                                                 \
      return launch_bwd<float, N>(student, teacher, center, row_stats, grad_out, B, K, inv_student_temp,   \
                                  inv_teacher_temp, grad_student, st);                                     \
    if (dtype == LAFS_BF16)                                                                                \
      return launch_bwd<__nv_bfloat16, N>(student, teacher, center, row_stats, grad_out, B, K,             \
                                          inv_student_temp, inv_teacher_temp, grad_student, st);           \
    return launch_bwd<__half, N>(student, teacher, center, row_stats, grad_out, B, K, inv_student_temp,    \
                                 inv_teacher_temp, grad_student, st);
  switch (ncrops) { LAFS_DINO_CROPS(LAFS_CASE) }
#undef LAFS_CASE
  return LAFS_ERR_ARG;
}

extern "C" int lafs_center_ema(const float* center, const float* colsum, float count, float momentum,
                               float one_minus_momentum, int K, float* center_out, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(center)) return brc;
  using namespace lafs;
  LAFS_REQUIRE(center && colsum && center_out && K > 0, LAFS_ERR_ARG, "lafs_center_ema: bad argument");
  launch_pdl((center_ema_kernel), dim3((K + 255) / 256), dim3(256), (size_t)(0), (cudaStream_t)stream, center, colsum, count, momentum,
                                                                      one_minus_momentum, K, center_out);
  return check_launch("lafs_center_ema");
}

extern "C" int lafs_colsum(const void* x, int rows, int K, int dtype, float* out, void* workspace,
                           size_t workspace_bytes, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(x)) return brc;
  using namespace lafs;
  LAFS_REQUIRE(x && out && workspace && rows > 0 && K > 0, LAFS_ERR_ARG, "lafs_colsum: bad argument");
  LAFS_REQUIRE(dtype >= 0 && dtype <= 2, LAFS_ERR_ARG, "lafs_colsum: dtype=%d", dtype);
  const int vec = dtype == LAFS_F32 ? 4 : 8;
  LAFS_REQUIRE(K % vec == 0 && ((uintptr_t)x & 15u) == 0, LAFS_ERR_ARG, "lafs_colsum: K %% %d != 0 or misaligned", vec);
  const int ngroups = rows < kMaxGroups ? rows : kMaxGroups;
  LAFS_REQUIRE(workspace_bytes >= (size_t)ngroups * K * sizeof(float), LAFS_ERR_WORKSPACE,
               "lafs_colsum: workspace too small (need %zu)", (size_t)ngroups * K * sizeof(float));
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((K / vec + kDinoThreads - 1) / kDinoThreads, ngroups);
  float* part = (float*)workspace;
  if (dtype == LAFS_F32) launch_pdl((colsum_partial_kernel<float>), dim3(grid), dim3(kDinoThreads), (size_t)(0), st, (const float*)x, rows, K, ngroups, part);
  else if (dtype == LAFS_BF16) launch_pdl((colsum_partial_kernel<__nv_bfloat16>), dim3(grid), dim3(kDinoThreads), (size_t)(0), st, (const __nv_bfloat16*)x, rows, K, ngroups, part);
  else launch_pdl((colsum_partial_kernel<__half>), dim3(grid), dim3(kDinoThreads), (size_t)(0), st, (const __half*)x, rows, K, ngroups, part);
  launch_pdl((colsum_combine_kernel), dim3((K + 255) / 256), dim3(256), (size_t)(0), st, part, ngroups, K, out);
  return check_launch("lafs_colsum");
}
