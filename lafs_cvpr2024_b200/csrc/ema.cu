// (3) Multi-tensor teacher EMA: one launch for every parameter tensor.
// Replaces the python loop lafs_train.py:610-613 (441 launches, 28 B/param) with a single
// pointer-table kernel moving the minimal 12 B/param (read k, read q, write k).
#include "common.cuh"
#include "../../include/lafs_b200.h"

namespace lafs {

struct EmaChunk {
  float* k;        // teacher run (in place)
  const float* q;  // student run
  int64_t n;       // <= LAFS_EMA_CHUNK
};
static_assert(sizeof(EmaChunk) == 24, "table record is 24 bytes");

constexpr int kEmaThreads = 256;

// k = fl(fl(k*m) + fl(q*om)) : three separately rounded fp32 ops == Tensor.mul_ / mul / add_.
__device__ __forceinline__ float ema1(float k, float q, float m, float om) {
  return __fadd_rn(__fmul_rn(k, m), __fmul_rn(q, om));
}

// grid == nchunks: one chunk per CTA.  grid < nchunks: persistent CTAs stride over the chunks -- used
// to co-schedule the EMA next to a kernel that leaves HBM bandwidth unused (one 256-thread CTA per SM).
__global__ void __launch_bounds__(kEmaThreads)
ema_multi_kernel(const EmaChunk* __restrict__ table, int nchunks, float m, float om) {
  pdl_wait();
 for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
  const EmaChunk c = table[chunk];
  const int n = (int)c.n;
  float* __restrict__ k = c.k;
  const float* __restrict__ q = c.q;
  const bool aligned = ((((uintptr_t)k) | ((uintptr_t)q)) & 15u) == 0;
  if (aligned) {
    const int n4 = n >> 2;
    // 4 independent 128-bit loads per operand in flight per thread
    int i = threadIdx.x;
    for (; i + 3 * kEmaThreads < n4; i += 4 * kEmaThreads) {
      float4 kv[4], qv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) kv[u] = ld_stream_f4(reinterpret_cast<const float4*>(k) + i + u * kEmaThreads);
#pragma unroll
      for (int u = 0; u < 4; ++u) qv[u] = ld_stream_f4(reinterpret_cast<const float4*>(q) + i + u * kEmaThreads);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float4 r;
        r.x = ema1(kv[u].x, qv[u].x, m, om);
        r.y = ema1(kv[u].y, qv[u].y, m, om);
        r.z = ema1(kv[u].z, qv[u].z, m, om);
        r.w = ema1(kv[u].w, qv[u].w, m, om);
        st_stream_f4(reinterpret_cast<float4*>(k) + i + u * kEmaThreads, r);
      }
    }
    for (; i < n4; i += kEmaThreads) {
      float4 kv = ld_stream_f4(reinterpret_cast<const float4*>(k) + i);
      float4 qv = ld_stream_f4(reinterpret_cast<const float4*>(q) + i);
      float4 r;
      r.x = ema1(kv.x, qv.x, m, om);
      r.y = ema1(kv.y, qv.y, m, om);
      r.z = ema1(kv.z, qv.z, m, om);
      r.w = ema1(kv.w, qv.w, m, om);
      st_stream_f4(reinterpret_cast<float4*>(k) + i, r);
    }
    for (int j = (n4 << 2) + threadIdx.x; j < n; j += kEmaThreads) k[j] = ema1(k[j], q[j], m, om);
  } else {
    for (int j = threadIdx.x; j < n; j += kEmaThreads) k[j] = ema1(k[j], q[j], m, om);
  }
 }
}

}  // namespace lafs

extern "C" int lafs_ema_multi(const void* table, int nchunks, float m, float one_minus_m, int max_ctas,
                              lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(table)) return brc;
  using namespace lafs;
  if (nchunks == 0) return LAFS_OK;
  LAFS_REQUIRE(table != nullptr && nchunks > 0, LAFS_ERR_ARG, "lafs_ema_multi: null table or nchunks<0");
  const int grid = (max_ctas > 0 && max_ctas < nchunks) ? max_ctas : nchunks;
  launch_pdl((ema_multi_kernel), dim3(grid), dim3(kEmaThreads), (size_t)(0), (cudaStream_t)stream, 
      reinterpret_cast<const EmaChunk*>(table), nchunks, m, one_minus_m);
  return check_launch("lafs_ema_multi");
}
