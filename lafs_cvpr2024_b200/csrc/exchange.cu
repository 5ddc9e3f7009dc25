// (4e) Class-sharded margin head: the two exchange steps over NVLink peer memory instead of NCCL.
//
// The head is class-parallel (rank r owns torch.chunk(weight, R)[r], ViT_face.py:56).  Per step the
// ranks exchange (a) the per-row softmax statistics (max, sum-exp, target logits: [B,4] fp32 per rank)
// and (b) the partial embedding gradient dE_hat [B,D] fp32, which has to be summed over the ranks.
// NCCL needs ~10-20 us per call for these few-KB / 1 MB messages; here each exchange is ONE kernel that
// moves the data with plain loads / stores through the peers' mapped memory (symmetric buffers from
// torch.distributed._symmetric_memory: same layout on every rank, base pointers in `peer_base`) and
// synchronises with release / acquire flags at system scope:
//   xchg_stats_kernel     put my [B,4] record array into every peer's slot -> flag -> wait for all
//                         peers' flags -> merge the R records per row (fixed rank order, so every rank
//                         computes bit-identical statistics).  Replaces all_gather + head_merge.
//   xchg_allreduce_kernel two-shot all-reduce: barrier, rank r sums slice r of dE_hat over all ranks
//                         (peer loads, fixed order), stores the sum into every rank's output (peer
//                         stores), barrier.  Replaces all_reduce(SUM); every rank gets identical bits.
// Flags carry a call counter (epoch) kept on the device, so the kernels can be replayed from a CUDA
// graph; slots are double buffered by epoch parity.  Spin loops are bounded: a dead peer turns into an
// error flag, not a hung GPU.
#include "common.cuh"
#include "../../include/lafs_b200.h"

namespace lafs {

constexpr int kXMaxRanks = 8;      // one NVSwitch domain
constexpr int kXBlocks = 16;       // flag / counter slots per kind (only slot 0 is used by the current kernels)
constexpr int kXReduceCtas = 64;   // CTAs of the all-reduce data phase
constexpr int kXThreads = 256;
constexpr long long kXSpinLimit = 1LL << 24;

// layout of the symmetric buffer (bytes); identical on every rank
struct XLayout {
  size_t off_flags;      // uint32 [3 kinds][kXBlocks][kXMaxRanks]
  size_t off_counters;   // uint32 [1 + kXBlocks] call counters (read/written by the owning rank only) + error flag
  size_t off_slots;      // float4 [2 parities][R][B]
  size_t off_in;         // float  [B][D]   this rank's partial dE_hat
  size_t off_out;        // float  [B][D]   the sum over ranks
  size_t total;
};
__host__ __device__ inline XLayout xlayout(int world, int B, int D) {
  XLayout l;
  size_t o = 0;
  l.off_flags = o;    o += (size_t)3 * kXBlocks * kXMaxRanks * 4;
  l.off_counters = o; o += 256;
  o = (o + 255) & ~(size_t)255;
  l.off_slots = o;    o += (size_t)2 * world * B * 16;
  o = (o + 255) & ~(size_t)255;
  l.off_in = o;       o += (size_t)B * D * 4;
  o = (o + 255) & ~(size_t)255;
  l.off_out = o;      o += (size_t)B * D * 4;
  l.total = (o + 255) & ~(size_t)255;
  return l;
}

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_f4(const float4* p) {   // never served from a stale L1 line
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// Threads 0..world-1 of the CTA: tell every peer "block `blk` of rank `rank` reached `epoch`" and wait
// until every peer said the same.  All prior writes of the CTA (ordered before this call by a
// __syncthreads) are made visible system-wide first.
__device__ __forceinline__ void xbarrier(const uint64_t* __restrict__ peer_base, size_t off_flags, int kind, int blk,
                                         int rank, int world, unsigned epoch, unsigned* err) {
  if ((int)threadIdx.x < world) {
    const int q = threadIdx.x;
    __threadfence_system();
    unsigned* theirs = reinterpret_cast<unsigned*>(peer_base[q] + off_flags) + ((size_t)kind * kXBlocks + blk) * kXMaxRanks + rank;
    st_release_sys(theirs, epoch);
    const unsigned* mine = reinterpret_cast<const unsigned*>(peer_base[rank] + off_flags) + ((size_t)kind * kXBlocks + blk) * kXMaxRanks + q;
    long long spins = 0;
    while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
      if (++spins > kXSpinLimit) { *err = 1u; break; }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kXThreads)
xchg_stats_kernel(const uint64_t* __restrict__ peer_base, int rank, int world, int B, XLayout l,
                  const float* __restrict__ local_stats, float* __restrict__ merged) {
  pdl_wait();
  __shared__ unsigned s_epoch;
  char* my = reinterpret_cast<char*>(peer_base[rank]);
  unsigned* counters = reinterpret_cast<unsigned*>(my + l.off_counters);
  if (threadIdx.x == 0) s_epoch = counters[0] + 1u;
  __syncthreads();
  const unsigned epoch = s_epoch;
  const int par = (int)(epoch & 1u);
  // put: my [B] records into slot (parity, rank) of every rank (my own included)
  const float4* src = reinterpret_cast<const float4*>(local_stats);
  for (int p = 0; p < world; ++p) {
    float4* dst = reinterpret_cast<float4*>(peer_base[p] + l.off_slots) + ((size_t)par * world + rank) * B;
    for (int b = threadIdx.x; b < B; b += kXThreads) dst[b] = src[b];
  }
  __syncthreads();
  xbarrier(peer_base, l.off_flags, 0, 0, rank, world, epoch, counters + 1 + kXBlocks);
  // merge the R records of every row in rank order (same arithmetic as head_merge_kernel)
  const float4* slots = reinterpret_cast<const float4*>(my + l.off_slots) + (size_t)par * world * B;
  for (int b = threadIdx.x; b < B; b += kXThreads) {
    float m = -INFINITY, s = 0.f, ta = 0.f, tb = 0.f;
    for (int q = 0; q < world; ++q) {
      const float4 v = ld_relaxed_sys_f4(slots + (size_t)q * B + b);
      if (v.x == -INFINITY) continue;
      const float mn = fmaxf(m, v.x);
      s = s * ex2(m - mn) + v.y * ex2(v.x - mn);
      m = mn;
      ta += v.z;
      tb += v.w;
    }
    reinterpret_cast<float4*>(merged)[b] = make_float4(m, s, ta, tb);
  }
  if (threadIdx.x == 0) counters[0] = epoch;
}

// Cross-GPU synchronisation is done by ONE CTA per phase (a system-scope round trip costs microseconds):
// CTA 0 runs the entry barrier and releases the other CTAs through a local flag; the CTA that finishes
// its stores last runs the exit barrier.  The data phase is spread over many CTAs with all peer loads
// of a thread issued before the first use (a peer load is ~2 us of latency).
__global__ void __launch_bounds__(kXThreads)
xchg_allreduce_kernel(const uint64_t* __restrict__ peer_base, int rank, int world, XLayout l, int n4) {
  pdl_wait();
  __shared__ unsigned s_epoch;
  __shared__ int s_last;
  char* my = reinterpret_cast<char*>(peer_base[rank]);
  unsigned* counters = reinterpret_cast<unsigned*>(my + l.off_counters);
  unsigned* err = counters + 1 + kXBlocks;
  unsigned* go = counters + 2;          // local: epoch whose entry barrier has completed
  unsigned* done = counters + 3;        // local: CTAs that finished their stores in this call
  if (threadIdx.x == 0) s_epoch = counters[1] + 1u;     // counters[1] is advanced by the last CTA only
  __syncthreads();
  const unsigned epoch = s_epoch;
  if (blockIdx.x == 0) {
    // every rank's partial (written by its preceding kernel) is complete
    xbarrier(peer_base, l.off_flags, 1, 0, rank, world, epoch, err);
    if (threadIdx.x == 0) {
      __threadfence();
      asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(go), "r"(epoch) : "memory");
    }
  } else {
    if (threadIdx.x == 0) {
      unsigned v;
      long long spins = 0;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(go) : "memory");
      } while ((int)(v - epoch) < 0 && ++spins < kXSpinLimit);
    }
    __syncthreads();
  }
  // slice of this rank, spread over the CTAs
  const int per_rank = (n4 + world - 1) / world;
  const int r_lo = rank * per_rank, r_hi = min(n4, r_lo + per_rank);
  constexpr int U = 4;
  for (int base = r_lo + ((int)blockIdx.x * kXThreads + (int)threadIdx.x) * U; base < r_hi;
       base += (int)gridDim.x * kXThreads * U) {
    float4 v[U][kXMaxRanks];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int q = 0; q < kXMaxRanks; ++q)
        if (q < world && base + u < r_hi)
          v[u][q] = ld_relaxed_sys_f4(reinterpret_cast<const float4*>(peer_base[q] + l.off_in) + base + u);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u < r_hi) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < kXMaxRanks; ++q)       // fixed order: bit-identical result on every rank
          if (q < world) { acc.x += v[u][q].x; acc.y += v[u][q].y; acc.z += v[u][q].z; acc.w += v[u][q].w; }
        for (int p = 0; p < world; ++p) reinterpret_cast<float4*>(peer_base[p] + l.off_out)[base + u] = acc;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    s_last = atomicAdd(done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last) {
    // every CTA of this rank has stored its part; tell the peers and wait until their slices have landed here
    xbarrier(peer_base, l.off_flags, 2, 0, rank, world, epoch, err);
    if (threadIdx.x == 0) { *done = 0u; counters[1] = epoch; }
  }
}

}  // namespace lafs

using namespace lafs;

/* out[5] = byte offsets of {flags, slots, partial dE_hat (input of the all-reduce), summed dE_hat, error flag} */
extern "C" size_t lafs_xchg_bytes(int world, int B, int D, size_t* offsets) {
  if (world < 1 || world > kXMaxRanks || B <= 0 || D <= 0) return 0;
  const XLayout l = xlayout(world, B, D);
  if (offsets != nullptr) {
    offsets[0] = l.off_flags; offsets[1] = l.off_slots; offsets[2] = l.off_in; offsets[3] = l.off_out;
    offsets[4] = l.off_counters + (size_t)(1 + kXBlocks) * 4;
  }
  return l.total;
}

static int xchg_check(const void* peer_base, int rank, int world, int B, int D, const char* who) {
  LAFS_REQUIRE(peer_base != nullptr, LAFS_ERR_ARG, "%s: null peer table", who);
  LAFS_REQUIRE(world >= 1 && world <= kXMaxRanks && rank >= 0 && rank < world, LAFS_ERR_ARG, "%s: rank=%d world=%d (<= %d ranks)",
               who, rank, world, kXMaxRanks);
  LAFS_REQUIRE(B > 0 && D > 0 && D % 4 == 0, LAFS_ERR_ARG, "%s: B=%d D=%d", who, B, D);
  return LAFS_OK;
}

/* merged[b] = merge over ranks of the per-row (max2, sum-exp, z_a, z_b) records; local_stats [B,4] fp32 is this
 * rank's record array (lafs_head_fwd's output).  peer_base: DEVICE array of `world` base addresses of the
 * symmetric buffer (lafs_xchg_bytes bytes each, zero-filled once before the first call). */
extern "C" int lafs_xchg_stats(const void* peer_base, int rank, int world, int B, int D, const float* local_stats,
                               float* merged, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(local_stats)) return brc;
  int rc = xchg_check(peer_base, rank, world, B, D, "lafs_xchg_stats");
  if (rc) return rc;
  LAFS_REQUIRE(local_stats && merged && ((((uintptr_t)local_stats | (uintptr_t)merged) & 15u) == 0), LAFS_ERR_ARG,
               "lafs_xchg_stats: null or misaligned statistics");
  launch_pdl((xchg_stats_kernel), dim3(1), dim3(kXThreads), (size_t)(0), (cudaStream_t)stream, (const uint64_t*)peer_base, rank, world, B,
                                                               xlayout(world, B, D), local_stats, merged);
  return check_launch("lafs_xchg_stats");
}

/* sum over ranks of the [B,D] fp32 array at offset[2] of every rank's buffer -> offset[3] of every rank's buffer */
extern "C" int lafs_xchg_allreduce(const void* peer_base, int rank, int world, int B, int D, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(peer_base)) return brc;
  int rc = xchg_check(peer_base, rank, world, B, D, "lafs_xchg_allreduce");
  if (rc) return rc;
  const long long n4 = (long long)B * D / 4;
  launch_pdl((xchg_allreduce_kernel), dim3(kXReduceCtas), dim3(kXThreads), (size_t)(0), (cudaStream_t)stream, (const uint64_t*)peer_base, rank, world,
                                                                         xlayout(world, B, D), (int)n4);
  return check_launch("lafs_xchg_allreduce");
}
