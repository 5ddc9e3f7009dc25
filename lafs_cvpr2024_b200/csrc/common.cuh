// Shared device/host helpers for liblafs_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "liblafs_b200 is written for sm_100a (B200) only"
#endif

namespace lafs {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- error plumbing (no exceptions cross the C ABI) --------------------------------------
void set_last_error(const char* fmt, ...);
int check_launch(const char* what);  // cudaGetLastError -> 0 or negative code, records message
// Makes the device that owns `device_ptr` current for THIS library's runtime on the calling thread.
// The library links its own (static) CUDA runtime, whose per-thread device defaults to 0; PyTorch's
// autograd worker threads set their device lazily, so every entry point binds explicitly.
int bind_device_of(const void* device_ptr);

#define LAFS_REQUIRE(cond, code, ...)                 \
  do {                                                \
    if (!(cond)) {                                    \
      ::lafs::set_last_error(__VA_ARGS__);            \
      return (code);                                  \
    }                                                 \
  } while (0)

enum : int {
  LAFS_OK = 0,
  LAFS_ERR_ARG = -1,       // bad argument (null pointer, unsupported size / dtype, misalignment)
  LAFS_ERR_CUDA = -2,      // a CUDA runtime call or launch failed (see lafs_last_error_string)
  LAFS_ERR_WORKSPACE = -3  // workspace too small
};

enum : int { LAFS_F32 = 0, LAFS_BF16 = 1, LAFS_F16 = 2, LAFS_U8 = 3 };

// ---- programmatic dependent launch (PDL) ------------------------------------------------------------------
// Every kernel of the library starts with pdl_wait() and is launched with the programmatic-stream-serialisation
// attribute: the next kernel of a stream / graph is scheduled while its predecessor drains and blocks in
// griddepcontrol.wait until the predecessor has completed and flushed its writes -- the same ordering as a plain
// launch without the launch gap (measured on B200, round 2: 83.3 -> 80.3 us on the two-kernel DINO forward).
// LAFS_PDL=0 launches without the attribute (pdl_wait() is then a no-op).
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif

// ---- device helpers ------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// float max over the warp: sm_100a reduces fp32 natively (one CREDUX.MAX.F32, result in a uniform
// register) -- no order-preserving integer key needed.
__device__ __forceinline__ float warp_max_redux(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

// Warp sum of values known to lie in [0, 256] per lane-group total (softmax partial sums relative to
// the slice maximum: every term is <= 1 and a 256-column slice holds at most 256 of them) with ONE
// integer REDUX: 2^-22 fixed point, i.e. <= 2^-17 absolute error on a sum that is >= 1.  Integer
// addition is associative, so the result is also order-independent.
__device__ __forceinline__ float warp_sum_unit_terms(float v) {
  int i = __float2int_rn(v * 4194304.0f);
  i = __reduce_add_sync(0xffffffffu, i);
  return (float)i * (1.0f / 4194304.0f);
}

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}

// streaming 128-bit accesses (data touched once: keep it out of L1)
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float4 ld_stream_f4(const void* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream_u4(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream_f4(void* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// 256-bit store (sm_100: STG.256): one instruction covers a 32-byte sector per lane.  p must be
// 32-byte aligned.
__device__ __forceinline__ void st_global_256(void* p, uint4 a, uint4 b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               :: "l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}

// 16-bit pair unpack / pack.  T = __nv_bfloat16 or __half.
template <typename T> struct Half2Ops;
template <> struct Half2Ops<__nv_bfloat16> {
  static __device__ __forceinline__ float lo(uint32_t u) { return __uint_as_float(u << 16); }
  static __device__ __forceinline__ float hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};
template <> struct Half2Ops<__half> {
  static __device__ __forceinline__ float lo(uint32_t u) {
    return __half2float(__ushort_as_half((unsigned short)(u & 0xffffu)));
  }
  static __device__ __forceinline__ float hi(uint32_t u) {
    return __half2float(__ushort_as_half((unsigned short)(u >> 16)));
  }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
};

}  // namespace lafs
