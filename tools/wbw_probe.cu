// Pure-write / read / copy bandwidth probe for B200 (development tool, not product code).
// Answers: is the "3.9 TB/s pure-write ceiling" measured in round 1 with SIMT stores a property
// of the part, or of the store instruction stream?  Variants:
//   simt128     st.global.v4 per lane, grid-stride
//   simt128na   st.global.L1::no_allocate.v4
//   simt256     st.global.v8 (STG.256)
//   tma         cp.async.bulk.global.shared::cta from a constant shared-memory tile (UBLKCP / bulk store)
//   memset      cudaMemsetAsync
//   read        ld.global.nc.v4 sum
//   copy        ld + st
//   rows64      the patch-embed epilogue pattern: a warp store covers 64 contiguous bytes of a 1536-byte row
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/wbw_probe tools/wbw_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__global__ void k_simt128(uint4* p, size_t n) {
  const uint4 v = make_uint4(1, 2, 3, 4);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void k_simt128na(uint4* p, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "r"(1), "r"(2), "r"(3), "r"(4) : "memory");
}
__global__ void k_simt256(uint4* p, size_t n) {   // n counts 16-byte units; each thread writes 32 B
  for (size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 2; i < n; i += (size_t)gridDim.x * blockDim.x * 2)
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p + i), "r"(1), "r"(2), "r"(3), "r"(4), "r"(5), "r"(6), "r"(7), "r"(8) : "memory");
}
__global__ void k_read(const uint4* p, size_t n, unsigned* out) {
  unsigned acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p + i));
    acc += r.x ^ r.y ^ r.z ^ r.w;
  }
  if (acc == 0x12345) *out = acc;
}
__global__ void k_copy(const uint4* s, uint4* d, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] = s[i];
}
// bulk (TMA) stores from shared memory: CHUNK bytes per instruction, one issuing thread per CTA
template <int CHUNK>
__global__ void k_tma(uint8_t* p, size_t nbytes) {
  extern __shared__ __align__(128) uint8_t sm[];
  for (int i = threadIdx.x; i < CHUNK / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(sm);
    int inflight = 0;
    for (size_t off = (size_t)blockIdx.x * CHUNK; off + CHUNK <= nbytes; off += (size_t)gridDim.x * CHUNK) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + off), "r"(src), "r"(CHUNK) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (++inflight >= 8) { asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory"); }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
// patch-embed epilogue pattern: lane = feature (2 B), 32 tokens per thread at `dim` stride
__global__ void k_rows64(uint16_t* p, int ntok_total, int dim) {
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int pieces_per_tokblock = dim / 32;
  const long long total = (long long)(ntok_total / 32) * pieces_per_tokblock;
  for (long long it = warp_global; it < total; it += nwarps) {
    const int tb = (int)(it / pieces_per_tokblock), pc = (int)(it % pieces_per_tokblock);
    uint16_t* q = p + (size_t)tb * 32 * dim + pc * 32 + lane;
#pragma unroll
    for (int j = 0; j < 32; ++j) q[(size_t)j * dim] = (uint16_t)j;
  }
}
// same bytes, but each thread owns 32 consecutive features of one token (64 B): 2 x STG.256
__global__ void k_rows64_t(uint16_t* p, int ntok_total, int dim) {
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long nthreads = (long long)gridDim.x * blockDim.x;
  const int pieces = dim / 32;
  const long long total = (long long)ntok_total * pieces;
  for (long long it = tid; it < total; it += nthreads) {
    const long long tok = (it / (32 * pieces)) * 32 + (it % 32);
    const int pc = (int)((it / 32) % pieces);
    uint16_t* q = p + (size_t)tok * dim + pc * 32;
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(q), "r"(1), "r"(2), "r"(3), "r"(4), "r"(5), "r"(6), "r"(7), "r"(8) : "memory");
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(q + 16), "r"(1), "r"(2), "r"(3), "r"(4), "r"(5), "r"(6), "r"(7), "r"(8) : "memory");
  }
}

template <typename F>
static float timeit(F f, int reps = 10) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) {
    CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main() {
  const size_t bytes = (size_t)2 << 30;   // 2 GiB: far larger than L2
  uint8_t *a, *b; unsigned* out;
  CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&out, 4));
  CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 1, bytes));
  const size_t n16 = bytes / 16;
  auto rep = [&](const char* name, float ms, double nb) { printf("%-28s %8.3f ms  %7.0f GB/s\n", name, ms, nb / ms / 1e6); fflush(stdout); };
  for (int cps : {4, 8, 16}) {
    const int grid = 148 * cps;
    char nm[64];
    snprintf(nm, 64, "simt128 %d CTA/SM", cps); rep(nm, timeit([&] { k_simt128<<<grid, 256>>>((uint4*)a, n16); }), (double)bytes);
    snprintf(nm, 64, "simt128na %d CTA/SM", cps); rep(nm, timeit([&] { k_simt128na<<<grid, 256>>>((uint4*)a, n16); }), (double)bytes);
    snprintf(nm, 64, "simt256 %d CTA/SM", cps); rep(nm, timeit([&] { k_simt256<<<grid, 256>>>((uint4*)a, n16); }), (double)bytes);
  }
  rep("memset", timeit([&] { CK(cudaMemsetAsync(a, 3, bytes)); }), (double)bytes);
  for (int cps : {1, 2, 4}) {
    char nm[64];
    CK(cudaFuncSetAttribute(k_tma<32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    snprintf(nm, 64, "tma 32KB x%d CTA/SM", cps); rep(nm, timeit([&] { k_tma<32768><<<148 * cps, 128, 32768>>>(a, bytes); }), (double)bytes);
    snprintf(nm, 64, "tma 8KB x%d CTA/SM", cps); rep(nm, timeit([&] { k_tma<8192><<<148 * cps, 128, 8192>>>(a, bytes); }), (double)bytes);
    snprintf(nm, 64, "tma 2KB x%d CTA/SM", cps); rep(nm, timeit([&] { k_tma<2048><<<148 * cps, 128, 2048>>>(a, bytes); }), (double)bytes);
  }
  rep("read 8 CTA/SM", timeit([&] { k_read<<<148 * 8, 256>>>((const uint4*)a, n16, out); }), (double)bytes);
  rep("copy 8 CTA/SM (r+w)", timeit([&] { k_copy<<<148 * 8, 256>>>((const uint4*)a, (uint4*)b, n16); }), 2.0 * bytes);
  rep("cudaMemcpy D2D (r+w)", timeit([&] { CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice)); }), 2.0 * bytes);
  // the patch-embed token tensor: 100352 + 100352 + 36864 token rows x 768 bf16 = 365 MB
  const int ntok = (100352 * 2 + 36864), dim = 768;
  const double tb = (double)ntok * dim * 2;
  rep("rows64 (epilogue now) x4", timeit([&] { k_rows64<<<148 * 4, 256>>>((uint16_t*)a, ntok, dim); }), tb);
  rep("rows64 (epilogue now) x8", timeit([&] { k_rows64<<<148 * 8, 256>>>((uint16_t*)a, ntok, dim); }), tb);
  rep("rows64 thread-row 256b x8", timeit([&] { k_rows64_t<<<148 * 8, 256>>>((uint16_t*)a, ntok, dim); }), tb);
  rep("simt128 365MB", timeit([&] { k_simt128<<<148 * 8, 256>>>((uint4*)a, (size_t)(tb / 16)); }), tb);
  CK(cudaFuncSetAttribute(k_tma<8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192));
  rep("tma 8KB 365MB", timeit([&] { k_tma<8192><<<148 * 2, 128, 8192>>>(a, (size_t)tb); }), tb);
  return 0;
}
