// (1b) Fused landmark patch gather -> patch_to_embedding on tcgen05 tensor cores.
// Replaces extract_patches_pytorch_gridsample + einops rearrange + nn.Linear(192, dim)
// (face_pre_pro/ViT_face.py:1615-1656, lafs_train.py:538, ViT_face.py:760-761): the fp32 mosaic,
// the re-laid-out token tensor and (for the two global SSL views) the second gather for the
// teacher are never written to HBM -- one pass reads each image once and writes only the
// embedded tokens of up to two models (student + teacher) that share the gathered patches.
//
// One persistent CTA per SM walks over faces.  Per face:
//   plane producer : cp.async.bulk of the three 112x112 channel planes into a 2-slot ring
//   gather warps   : bilinear 8x8 patches from the staged plane -> bf16 token tile in the UMMA
//                    K-major/128B-swizzle layout (tokens x 192), K order = c*64 + j*8 + i
//   W producer     : TMA of [128 x 64] bf16 weight chunks (weights pre-permuted to that K order)
//   UMMA issuer    : D[128 dims x N tokens] += Wchunk[128 x 192] * Tok[N x 192]^T   (tokens on the
//                    N side: N = 208 for 196 landmarks, 48 for 36 -- no 128-row padding waste)
//   epilogue warps : TMEM -> +bias -> bf16 -> out[model][face, token, dim]
//
// Two input formats:
//   fp32 planes (the reference's tensors).  100 KB of planes leave room for ONE token tile, so
//     the gather of face f+1 waits for the UMMAs of face f.
//   uint8 planes (what the data loader decodes; ToTensor + Normalize = a*u8 + b is applied to the
//     interpolated value, zero padding becomes the raw value -b/a).  25 KB of planes leave room
//     for TWO token tiles: gather(f+1) overlaps UMMA/epilogue(f), and HBM/PCIe image bytes drop 4x.
//
// The bf16 path does not need the reference's fp32 coordinate round trip: sample
// positions are theta + (idx - 4.5) directly; the error (<1e-5 px) is far below bf16 rounding.
// The stand-alone fp32 gather (gather.cu) keeps the exact sequence for the 1e-5 parity clause.
#include <stdlib.h>
#include "umma.cuh"
#include "../../include/lafs_b200.h"

namespace lafs {
using namespace umma;

namespace pe {
constexpr int kH = 112, kW = 112, kC = 3;
constexpr int kFeat = 192;                         // 3 * 8 * 8
constexpr int kTokLd = 208;                        // row pitch of the saved-token tensor: 192 features + a ones column
                                                   // (bias gradient) + 15 zero columns (16-byte row alignment)
constexpr int kMaxTok = 208;                       // largest UMMA N (196 landmarks -> N_pad 208)
constexpr int kTokRows = 200;                      // token rows owned per chunk (25 x 8); a UMMA with
                                                   // N_pad = 208 also reads 8 rows of the NEXT region:
                                                   // harmless garbage columns that are never stored
constexpr int kTokChunkBytes = kTokRows * 128;     // one 64-feature chunk: 25,600 B (25 x 1024)
constexpr int kTokTileBytes = 3 * kTokChunkBytes;  // 76,800 B
constexpr int kWStageBytes = 128 * 128;            // [128 dims x 64 k] bf16
constexpr int kWStages = 3;
constexpr int kEpiWarps = 8;                       // two warps per TMEM lane quarter
constexpr int kScratchBytes = 0;
constexpr int kGatherWarps = 8;
constexpr int kGatherThreads = kGatherWarps * 32;
// warp roles
constexpr int kWarpPlane = 0, kWarpW = 1, kWarpMma = 2, kWarpEpi0 = 4, kWarpGather0 = kWarpEpi0 + kEpiWarps;
constexpr int kThreads = (kWarpGather0 + kGatherWarps) * 32;   // 20 warps

// (A variant with a 7-slot plane ring and one token tile for the 36-landmark views was measured on B200 in
//  round 2 and removed: 59.7 us against 53.3 us for this layout on 1024 x 36 uint8 views.)
template <typename InT>
struct Layout {
  static constexpr int kPlaneBytes = kH * kW * (int)sizeof(InT);          // 50,176 (fp32) / 12,544 (u8)
  static constexpr int kPlaneSlot = (kPlaneBytes + 1023) / 1024 * 1024;   // keeps later regions 1024-aligned
  static constexpr int kPlaneSlots = 2;
  static constexpr int kTokBufs = sizeof(InT) == 1 ? 2 : 1;
  static constexpr int kOffPlanes = 0;
  static constexpr int kOffTok = kPlaneSlots * kPlaneSlot;
  static constexpr int kOffW = kOffTok + kTokBufs * kTokTileBytes;
  static constexpr int kOffScratch = kOffW + kWStages * kWStageBytes;
  static constexpr int kOffBar = kOffScratch + kEpiWarps * kScratchBytes;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes = kOffBar + kBarBytes + 1024;           // + barriers + alignment slack
  static_assert(kOffTok % 1024 == 0 && kOffW % 1024 == 0, "UMMA tiles need 1024-byte alignment");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};
}  // namespace pe

struct EmbedParams {
  const void* imgs;         // [Bv, 3, 112, 112] fp32 or uint8
  const float* theta;       // [Bv, n, 2]
  const float* bias;        // [n_models * dim]
  void* out[2];             // per model: [Bv, n, dim] bf16 or fp32
  __nv_bfloat16* tok_out;   // optional: the gathered tokens, bf16 [Bv*n, kTokLd] in the kernel's K order (c*64+j*8+i),
                            //   kept for the weight-gradient GEMM of the training path (columns >= 192 are the caller's)
  int Bv, n, n_pad, dim, n_models;
  int gfaces;               // faces whose tokens share one tile / one pass over the weights (G*n <= 200)
  int ngroups;              // ceil(Bv / gfaces)
  int nfull;                // groups [0, nfull) are whole work units; each later group is split into
  int nunits;               //   two units (first / second half of the weight chunks) for load balance
  int mchunks;              // n_models * dim / 128
  float in_scale, in_shift; // normalised pixel = in_scale * raw + in_shift   (1, 0 for fp32 input)
  float pad_raw;            // raw value of a zero-padded (out-of-image) pixel = -in_shift / in_scale
  int debug;                // development ablations (env LAFS_PE_DEBUG): 1 = no stores, 2 = no gather math, 4 = no UMMA
  // sequence epilogue (ViT_face.py:762-768): out_m is [Bv, n+1, dim]; row 0 = cls_token + pos[0],
  // row 1+t = embedding[t] + pos[1+t], then dropout(p) -- the torch.cat / += / dropout passes never run
  int seq;                  // 0: plain [Bv, n, dim] embeddings
  const float* pos[2];      // per model: pos_embedding rows [>= n+1, dim] fp32
  const float* cls[2];      // per model: cls_token [dim] fp32
  float drop_p, drop_scale; // dropout probability and 1/(1-p)  (p = 0: off)
  uint32_t drop_seed;
};

// counter-based uniform in [0,1) for the fused dropout: a 32-bit mix of (seed, element index).  The mask is a
// function of (seed, model, output element) only, so it is reproducible and independent of the launch geometry;
// it cannot (and need not) reproduce torch's Philox stream -- the reference is itself stochastic here.
__device__ __forceinline__ float hash_uniform(uint32_t seed, uint32_t idx) {
  uint32_t x = idx * 0x9E3779B1u + seed;
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return (float)(x >> 8) * (1.0f / 16777216.0f);
}

__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void store_out(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_out(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float raw_pixel(const float* plane, int idx) { return plane[idx]; }
__device__ __forceinline__ float raw_pixel(const uint8_t* plane, int idx) { return (float)plane[idx]; }

// Bilinear gather of one staged channel plane into one 64-feature chunk of the token tile.
template <typename InT, int JB>
__device__ __forceinline__ void gather_plane(const InT* __restrict__ plane, uint8_t* __restrict__ tok,
                                             const float* __restrict__ th, int n, int row0, int gt, float a_in,
                                             float b_in, float pad, int debug, __nv_bfloat16* __restrict__ tok_g) {
  using namespace pe;
  constexpr int NB = 8 / JB;                               // items per token
  for (int item = gt; item < ((debug & 2) ? 0 : NB * n); item += kGatherThreads) {
    const int t = item / NB, h = item - t * NB;
    const float2 thv = __ldg(reinterpret_cast<const float2*>(th) + t);
    const float sx = thv.x - 4.5f, sy = thv.y - 4.5f + (float)(JB * h);
    const float fxf = floorf(sx), fyf = floorf(sy);
    const float wx = sx - fxf, wy = sy - fyf;
    const int x0 = (int)fminf(fmaxf(fxf, -16.f), 128.f);
    const int y0 = (int)fminf(fmaxf(fyf, -16.f), 128.f);
    const bool inside = (x0 >= 0) && (x0 + 8 < kW) && (y0 >= 0) && (y0 + JB < kH);
    float hprev[8];
#pragma unroll
    for (int r = 0; r < JB + 1; ++r) {
      float px[9];
      const int y = y0 + r;
      if (inside) {
        const int base = y * kW + x0;
        if constexpr (sizeof(InT) == 1) {
          // 9 consecutive bytes = 3 aligned 32-bit shared loads + funnel shifts instead of 9 byte loads (the
          // byte loads of 32 lanes at unrelated addresses were the kernel's bank-conflict source); rows are
          // 112 B, so y*kW is word aligned and the third word stays inside the row (x0 + 8 < 112)
          const uint32_t* wp = reinterpret_cast<const uint32_t*>(plane) + (base >> 2);
          const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
          const int sh = (base & 3) * 8;
          const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh), v2 = w2 >> sh;
          px[0] = (float)(v0 & 0xffu); px[1] = (float)((v0 >> 8) & 0xffu); px[2] = (float)((v0 >> 16) & 0xffu);
          px[3] = (float)(v0 >> 24);
          px[4] = (float)(v1 & 0xffu); px[5] = (float)((v1 >> 8) & 0xffu); px[6] = (float)((v1 >> 16) & 0xffu);
          px[7] = (float)(v1 >> 24);
          px[8] = (float)(v2 & 0xffu);
        } else {
#pragma unroll
          for (int q = 0; q < 9; ++q) px[q] = raw_pixel(plane, base + q);
        }
      } else {
        const bool yok = (y >= 0) && (y < kH);
        const int rowb = min(max(y, 0), kH - 1) * kW;
#pragma unroll
        for (int q = 0; q < 9; ++q) {
          const int x = x0 + q;
          const float v = raw_pixel(plane, rowb + min(max(x, 0), kW - 1));
          px[q] = (yok && x >= 0 && x < kW) ? v : pad;
        }
      }
      float hcur[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) hcur[i] = fmaf(wx, px[i + 1] - px[i], px[i]);
      if (r > 0) {
        const int j = JB * h + r - 1;
        uint4 pk;
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          o[i] = fmaf(wy, hcur[i] - hprev[i], hprev[i]);
          if (sizeof(InT) == 1) o[i] = fmaf(o[i], a_in, b_in);   // ToTensor + Normalize, after the lerp
        }
        pk.x = Half2Ops<__nv_bfloat16>::pack(o[0], o[1]);
        pk.y = Half2Ops<__nv_bfloat16>::pack(o[2], o[3]);
        pk.z = Half2Ops<__nv_bfloat16>::pack(o[4], o[5]);
        pk.w = Half2Ops<__nv_bfloat16>::pack(o[6], o[7]);
        *reinterpret_cast<uint4*>(tok + sw128_offset(row0 + t, j * 8)) = pk;
        // training path: the same 8 features also go to HBM (14 % of the output bytes) for the dW GEMM
        if (tok_g != nullptr) *reinterpret_cast<uint4*>(tok_g + (size_t)t * kTokLd + j * 8) = pk;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) hprev[i] = hcur[i];
    }
  }
}

// DIM > 0: compile-time embedding width (store offsets become immediates); DIM == 0: runtime p.dim
// Orientation: D[128 dims x N tokens], tokens on the UMMA N side (N = 208 for 196 landmarks: no 128-row padding).
// The transposed form (tokens on the M side, a TMEM lane = a token, 64 contiguous output bytes per thread and
// 16x fewer store instructions) was built and measured on B200 in round 2 and removed: 137.6 us against 98.9 us
// (512 x 196 uint8 views, two models) -- 196 tokens need two M tiles (23 % more UMMA work), and the ablations
// (profiles/r02_pe_ablate.txt) show the store STREAM (308 MB: ~50 us at the part's write rate), not the store
// instruction count, is what the epilogue costs.
template <typename InT, typename OutT, int DIM>
__global__ void __launch_bounds__(pe::kThreads, 1)
gather_embed_kernel(const __grid_constant__ CUtensorMap tmap_w, const EmbedParams p) {
  pdl_wait();
  using namespace pe;
  using L = Layout<InT>;
  constexpr int NTB = L::kTokBufs;
  constexpr int NPS = L::kPlaneSlots;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* s_planes = smem + L::kOffPlanes;
  uint8_t* s_tok = smem + L::kOffTok;
  uint8_t* s_w = smem + L::kOffW;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* plane_full = bars;                   // NPS
  uint64_t* plane_empty = bars + NPS;            // NPS
  uint64_t* tok_full = bars + 2 * NPS;           // NTB * 3 (room for 6)
  uint64_t* tok_empty = bars + 2 * NPS + 6;      // NTB (room for 2)
  uint64_t* w_full = bars + 2 * NPS + 8;         // kWStages (<= 4)
  uint64_t* w_empty = bars + 2 * NPS + 12;       // kWStages
  uint64_t* acc_full = bars + 2 * NPS + 16;      // 2
  uint64_t* acc_empty = bars + 2 * NPS + 18;     // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NPS + 20);
  static_assert((2 * NPS + 21) * 8 <= L::kBarBytes, "barrier region");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tensormap(&tmap_w);
    for (int i = 0; i < NPS; ++i) { mbar_init(plane_full + i, 1); mbar_init(plane_empty + i, kGatherWarps); }
    for (int i = 0; i < NTB * 3; ++i) mbar_init(tok_full + i, kGatherWarps);
    for (int i = 0; i < NTB; ++i) mbar_init(tok_empty + i, 1);
    for (int i = 0; i < kWStages; ++i) { mbar_init(w_full + i, 1); mbar_init(w_empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, kEpiWarps); }
    fence_mbar_init();
  }
  if (warp == kWarpMma) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work unit = a GROUP of p.gfaces consecutive faces (1 for 196 landmarks, 5 for 36): their tokens
  // fill one tile, so every weight chunk fetched from L2 is used for all of them
  // The units of the last, partially filled round are half groups (same tokens, half of the weight
  // chunks each), so that round costs half a unit instead of a whole one.
  const int nunits_mine = (p.nunits - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  auto unit_group = [&](int u) { return u < p.nfull ? u : p.nfull + ((u - p.nfull) >> 1); };
  auto unit_mc0 = [&](int u) { return u < p.nfull ? 0 : ((u - p.nfull) & 1) * (p.mchunks >> 1); };
  auto unit_mcn = [&](int u) { return u < p.nfull ? p.mchunks : (p.mchunks >> 1); };
  auto group_faces = [&](int g) { return min(p.gfaces, p.Bv - g * p.gfaces); };
  auto group_ntok = [&](int g) { return group_faces(g) * p.n; };

  if (warp == kWarpPlane) {
    // ===================== image plane producer =====================
    if (lane == 0) {
      const InT* imgs = reinterpret_cast<const InT*>(p.imgs);
      // (An L2 prefetch of the next unit's planes -- cp.async.bulk.prefetch.L2 one work unit ahead of the
      //  ring's loads -- was measured and removed: the bulk loads did not hit the prefetched lines, DRAM
      //  reads rose from 155 to 217 MB on the 36-landmark launch and the kernel slowed from 64 to 74 us.)
      uint32_t cnt = 0;
      for (int gi = 0; gi < nunits_mine; ++gi) {
        const int g = unit_group(blockIdx.x + gi * gridDim.x);
        const int nf = group_faces(g);
        for (int ff = 0; ff < nf; ++ff) {
          const int f = g * p.gfaces + ff;
          for (int c = 0; c < kC; ++c, ++cnt) {
            const int slot = cnt % NPS;
            mbar_wait(plane_empty + slot, ((cnt / NPS) & 1) ^ 1);
            mbar_arrive_expect_tx(plane_full + slot, L::kPlaneBytes);
            bulk_load(s_planes + slot * L::kPlaneSlot, imgs + ((size_t)f * kC + c) * (kH * kW), L::kPlaneBytes,
                      plane_full + slot);
          }
        }
      }
    }
  } else if (warp == kWarpW) {
    // ===================== weight chunk producer =====================
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int gi = 0; gi < nunits_mine; ++gi) {
        const int u = blockIdx.x + gi * gridDim.x;
        const int mc0 = unit_mc0(u), mc1 = mc0 + unit_mcn(u);
        for (int mc = mc0; mc < mc1; ++mc)
          for (int c = 0; c < kC; ++c, ++cnt) {
            const int st = cnt % kWStages;
            mbar_wait(w_empty + st, ((cnt / kWStages) & 1) ^ 1);
            mbar_arrive_expect_tx(w_full + st, kWStageBytes);
            tma_load_2d(s_w + st * kWStageBytes, &tmap_w, w_full + st, c * 64, mc * 128);
          }
      }
    }
  } else if (warp == kWarpMma) {
    // ===================== UMMA issuer =====================
    if (lane == 0) {
      uint32_t wcnt = 0, acnt = 0;
      for (int fi = 0; fi < nunits_mine; ++fi) {
        const int u = blockIdx.x + fi * gridDim.x;
        const int tb = fi % NTB;
        const uint32_t tuse = (uint32_t)(fi / NTB);
        const int npad_g = (group_ntok(unit_group(u)) + 15) & ~15;
        const uint32_t idesc = make_idesc_bf16(128, npad_g);
        const int mc0 = unit_mc0(u), mc1 = mc0 + unit_mcn(u);
        for (int mc = mc0; mc < mc1; ++mc, ++acnt) {
          const int buf = acnt & 1;
          mbar_wait(acc_empty + buf, ((acnt >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256);
          for (int c = 0; c < kC; ++c, ++wcnt) {
            if (mc == mc0) { mbar_wait(tok_full + tb * 3 + c, tuse & 1); }
            const int st = wcnt % kWStages;
            mbar_wait(w_full + st, (wcnt / kWStages) & 1);
            tc_fence_after();
            const uint64_t dw = make_desc_k_sw128(smem_u32(s_w + st * kWStageBytes));
            const uint32_t tok_addr = smem_u32(s_tok + tb * kTokTileBytes + c * kTokChunkBytes);
            if (!(p.debug & 4)) {
              const uint64_t dt = make_desc_k_sw128(tok_addr);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                mma_f16_ss(d_tmem, desc_advance_k(dw, kk * 16), desc_advance_k(dt, kk * 16), idesc, (c | kk) != 0);
            }
            mma_commit(w_empty + st);
          }
          mma_commit(acc_full + buf);
        }
        mma_commit(tok_empty + tb);   // all UMMAs reading this face's token tile have completed
      }
    }
  } else if (warp >= kWarpEpi0 && warp < kWarpEpi0 + kEpiWarps) {
    // ===================== epilogue =====================
    const int quarter = warp & 3;
    const int half = (warp - kWarpEpi0) >> 2;
    const int dim = DIM > 0 ? DIM : p.dim;
    uint32_t acnt = 0;
    // lane = output feature (TMEM lane), registers = 32 consecutive tokens.  A warp-wide 2-byte
    // store covers 32 consecutive features of one token (64 contiguous bytes = 2 full sectors).
    // The two warps of a lane quarter take alternate 32-token pieces; the TMEM load of the next
    // piece is in flight while the current one is converted and stored.
    // (A shared-memory transpose to 16-byte stores was measured 25 % SLOWER: the extra STS/LDS
    //  traffic competes with the gather warps for the MIO pipe.)
    for (int fi = 0; fi < nunits_mine; ++fi) {
      const int u = blockIdx.x + fi * gridDim.x;
      const int g = unit_group(u);
      const int f = g * p.gfaces;                          // first face: the group's tokens are contiguous in `out`
      const int ntok = group_ntok(g), npad = (ntok + 15) & ~15;
      const int mc0 = unit_mc0(u), mc1 = mc0 + unit_mcn(u);
      for (int mc = mc0; mc < mc1; ++mc, ++acnt) {
        const int buf = acnt & 1;
        const int d = mc * 128 + quarter * 32 + lane;      // row of the stacked [n_models*dim] weight
        const int model = d >= dim ? 1 : 0, dd = d - model * dim;
        const float bias = __ldg(p.bias + d);
        OutT* dst = reinterpret_cast<OutT*>(model == 0 ? p.out[0] : p.out[1]) + (size_t)f * p.n * dim + dd;
        mbar_wait(acc_full + buf, (acnt >> 1) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t)(buf * 256) + ((uint32_t)(quarter * 32) << 16);
        // n_pad is a multiple of 16; pieces are 32 columns, the last one may be a 16-column tail
        auto issue_ld = [&](int t0, uint32_t (&v)[32]) {
          if (t0 + 32 <= npad) {
            tmem_ld_32x32b_x32(taddr + (uint32_t)t0, v);
          } else {
            uint32_t lo[16];
            tmem_ld_32x32b_x16(taddr + (uint32_t)t0, lo);
#pragma unroll
            for (int j = 0; j < 16; ++j) { v[j] = lo[j]; v[16 + j] = 0u; }
          }
        };
        auto emit = [&](int t0, const uint32_t (&v)[32]) {
          if (p.debug & 1) return;
          OutT* q = dst + (size_t)t0 * dim;
          if (t0 + 32 <= ntok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) store_out(q + (size_t)j * dim, __uint_as_float(v[j]) + bias);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (t0 + j < ntok) store_out(q + (size_t)j * dim, __uint_as_float(v[j]) + bias);
          }
        };
        // sequence form: token tt of the group = (face f + tt / n, landmark t = tt % n) -> output row
        // (f + tt/n)*(n+1) + 1 + t, plus pos_embedding[1 + t] and dropout; the thread that owns landmark 0 of a
        // face also writes that face's cls row
        auto emit_seq = [&](int t0, const uint32_t (&v)[32]) {
          if (p.debug & 1) return;
          const float* pos = model == 0 ? p.pos[0] : p.pos[1];
          OutT* base = reinterpret_cast<OutT*>(model == 0 ? p.out[0] : p.out[1]);
          const uint32_t mseed = p.drop_seed + (uint32_t)model * 0x85EBCA6Bu;
          int ff = t0 / p.n, t = t0 - ff * p.n;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            if (t0 + j < ntok) {
              const size_t row = (size_t)(f + ff) * (p.n + 1) + 1 + t;
              float val = __uint_as_float(v[j]) + bias + __ldg(pos + (size_t)(1 + t) * dim + dd);
              if (p.drop_p > 0.f)
                val = hash_uniform(mseed, (uint32_t)(row * dim + dd)) < p.drop_p ? 0.f : val * p.drop_scale;
              store_out(base + row * dim + dd, val);
              if (t == 0) {
                const float* cls = model == 0 ? p.cls[0] : p.cls[1];
                float cv = __ldg(cls + dd) + __ldg(pos + dd);
                if (p.drop_p > 0.f)
                  cv = hash_uniform(mseed, (uint32_t)((row - 1) * dim + dd)) < p.drop_p ? 0.f : cv * p.drop_scale;
                store_out(base + (row - 1) * dim + dd, cv);
              }
            }
            if (++t == p.n) { t = 0; ++ff; }
          }
        };
        // (software-pipelining the TMEM loads over two register sets bought nothing and cost 24
        //  registers per thread, which matter for co-residency with the EMA kernel)
        for (int t0 = half * 32; t0 < ntok; t0 += 64) {
          uint32_t v[32];
          issue_ld(t0, v);
          tmem_ld_wait();
          if (p.seq) emit_seq(t0, v);
          else emit(t0, v);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_relaxed(acc_empty + buf);   // no release: do not wait for the token stores
      }
    }
  } else if (warp >= kWarpGather0) {
    // ===================== gather warps =====================
    const int gt = threadIdx.x - kWarpGather0 * 32;        // 0..255
    const float a_in = p.in_scale, b_in = p.in_shift, pad = p.pad_raw;
    uint32_t pcnt = 0;
    for (int fi = 0; fi < nunits_mine; ++fi) {
      const int g = unit_group(blockIdx.x + fi * gridDim.x);
      const int nf = group_faces(g);
      const int tb = fi % NTB;
      const uint32_t tuse = (uint32_t)(fi / NTB);
      mbar_wait(tok_empty + tb, (tuse & 1) ^ 1);           // the UMMAs of the group that used this tile are done
     for (int ff = 0; ff < nf; ++ff) {
      const int f = g * p.gfaces + ff;
      const int row0 = ff * p.n;                           // first token row of this face inside the tile
      const float* th = p.theta + (size_t)f * p.n * 2;
      for (int c = 0; c < kC; ++c, ++pcnt) {
        const int slot = pcnt % NPS;
        mbar_wait(plane_full + slot, (pcnt / NPS) & 1);
        const InT* plane = reinterpret_cast<const InT*>(s_planes + slot * L::kPlaneSlot);
        uint8_t* tok = s_tok + tb * kTokTileBytes + c * kTokChunkBytes;
        // item = (token t, block h of JB output columns j): JB+1 pixel rows -> JB 16-byte stores.
        // JB = 4 (two items per token) keeps row reuse high when there are many tokens; with few
        // tokens per plane (36-landmark views) JB = 1 spreads the work over all gather threads.
        __nv_bfloat16* tok_g = p.tok_out != nullptr ? p.tok_out + (size_t)f * p.n * kTokLd + c * 64 : nullptr;
        if (p.n > 64) gather_plane<InT, 4>(plane, tok, th, p.n, row0, gt, a_in, b_in, pad, p.debug, tok_g);
        else          gather_plane<InT, 1>(plane, tok, th, p.n, row0, gt, a_in, b_in, pad, p.debug, tok_g);
        fence_proxy_async_smem();      // token stores -> visible to the UMMA (async proxy) reads
        __syncwarp();
        if (lane == 0) {
          if (ff == nf - 1) mbar_arrive(tok_full + tb * 3 + c);   // chunk c complete for every face of the group
          mbar_arrive(plane_empty + slot);
        }
      }
     }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// W fp32 [dim, 192] (feature f = (i*8+j)*3 + c) -> bf16 [dim, 192] with K order c*64 + j*8 + i
__global__ void embed_weight_prep_kernel(const float* __restrict__ w, const float* __restrict__ bias, int dim,
                                        __nv_bfloat16* __restrict__ out, float* __restrict__ bias_out) {
  pdl_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < dim && bias_out != nullptr) bias_out[idx] = bias != nullptr ? bias[idx] : 0.f;
  if (idx >= dim * pe::kFeat) return;
  const int d = idx / pe::kFeat, k = idx - d * pe::kFeat;
  const int c = k >> 6, j = (k >> 3) & 7, i = k & 7;
  out[idx] = __float2bfloat16_rn(w[(size_t)d * pe::kFeat + (i * 8 + j) * 3 + c]);
}

template <typename InT>
static int launch_embed(const CUtensorMap& tw, const EmbedParams& p, int out_dtype, int dim, cudaStream_t st) {
  void (*kern)(const CUtensorMap, const EmbedParams);
  if (out_dtype == LAFS_F32) kern = dim == 768 ? gather_embed_kernel<InT, float, 768> : gather_embed_kernel<InT, float, 0>;
  else kern = dim == 768 ? gather_embed_kernel<InT, __nv_bfloat16, 768>
            : dim == 384 ? gather_embed_kernel<InT, __nv_bfloat16, 384> : gather_embed_kernel<InT, __nv_bfloat16, 0>;
  constexpr int smem = pe::Layout<InT>::kSmemBytes;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  LAFS_REQUIRE(e == cudaSuccess, LAFS_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  const int grid = p.nunits < kNumSMs ? p.nunits : kNumSMs;
  launch_pdl((kern), dim3(grid), dim3(pe::kThreads), (size_t)(smem), st, tw, p);
  return check_launch("lafs_gather_embed_fwd");
}

}  // namespace lafs

using namespace lafs;

extern "C" int lafs_embed_weight_prep(const float* weight, const float* bias, int dim, void* out_bf16, float* bias_out,
                                      lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(weight)) return brc;
  LAFS_REQUIRE(weight && out_bf16 && dim > 0, LAFS_ERR_ARG, "lafs_embed_weight_prep: bad argument");
  const int total = dim * pe::kFeat;
  launch_pdl((embed_weight_prep_kernel), dim3((total + 255) / 256), dim3(256), (size_t)(0), (cudaStream_t)stream, weight, bias, dim, (__nv_bfloat16*)out_bf16, bias_out);
  return check_launch("lafs_embed_weight_prep");
}

extern "C" int lafs_gather_embed_fwd(const void* imgs, int in_dtype, float in_scale, float in_shift, const float* theta,
                                     const void* w_perm_bf16, const float* bias, void* out0, void* out1, int out_dtype,
                                     int Bv, int H, int W, int n, int dim, int n_models, lafs_stream_t stream) {
  return lafs_gather_embed_fwd_save(imgs, in_dtype, in_scale, in_shift, theta, w_perm_bf16, bias, out0, out1, out_dtype, Bv, H, W, n,
                                    dim, n_models, nullptr, stream);
}

extern "C" int lafs_gather_embed_fwd_save(const void* imgs, int in_dtype, float in_scale, float in_shift, const float* theta,
                                          const void* w_perm_bf16, const float* bias, void* out0, void* out1, int out_dtype,
                                          int Bv, int H, int W, int n, int dim, int n_models, void* tokens_perm_out,
                                          lafs_stream_t stream) {
  return lafs_gather_embed_seq_fwd(imgs, in_dtype, in_scale, in_shift, theta, w_perm_bf16, bias, out0, out1, out_dtype, Bv, H, W, n,
                                   dim, n_models, tokens_perm_out, nullptr, nullptr, nullptr, nullptr, 0.f, 0u, stream);
}

extern "C" int lafs_gather_embed_seq_fwd(const void* imgs, int in_dtype, float in_scale, float in_shift, const float* theta,
                                         const void* w_perm_bf16, const float* bias, void* out0, void* out1, int out_dtype,
                                         int Bv, int H, int W, int n, int dim, int n_models, void* tokens_perm_out,
                                         const float* pos0, const float* cls0, const float* pos1, const float* cls1,
                                         float drop_p, unsigned int drop_seed, lafs_stream_t stream) {
  if (int brc = lafs::bind_device_of(imgs)) return brc;
  if (Bv == 0) return LAFS_OK;
  LAFS_REQUIRE(imgs && theta && w_perm_bf16 && bias && out0, LAFS_ERR_ARG, "lafs_gather_embed_fwd: null pointer");
  LAFS_REQUIRE(in_dtype == LAFS_F32 || in_dtype == LAFS_U8, LAFS_ERR_ARG, "lafs_gather_embed_fwd: in_dtype=%d (fp32 or uint8)", in_dtype);
  LAFS_REQUIRE(out_dtype == LAFS_BF16 || out_dtype == LAFS_F32, LAFS_ERR_ARG, "lafs_gather_embed_fwd: out_dtype=%d (bf16 or fp32)", out_dtype);
  LAFS_REQUIRE(H == pe::kH && W == pe::kW, LAFS_ERR_ARG, "lafs_gather_embed_fwd: fused path is built for 112x112 faces, got %dx%d", H, W);
  LAFS_REQUIRE(n_models == 1 || (n_models == 2 && out1), LAFS_ERR_ARG, "lafs_gather_embed_fwd: n_models=%d", n_models);
  LAFS_REQUIRE(n > 0 && n <= pe::kMaxTok, LAFS_ERR_ARG, "lafs_gather_embed_fwd: n=%d outside [1,%d]", n, pe::kMaxTok);
  LAFS_REQUIRE(dim > 0 && dim % 128 == 0, LAFS_ERR_ARG, "lafs_gather_embed_fwd: dim=%d must be a multiple of 128", dim);
  LAFS_REQUIRE(((uintptr_t)imgs & 15u) == 0 && ((uintptr_t)theta & 7u) == 0, LAFS_ERR_ARG, "lafs_gather_embed_fwd: misaligned input");
  LAFS_REQUIRE(Bv > 0, LAFS_ERR_ARG, "lafs_gather_embed_fwd: Bv=%d", Bv);
  LAFS_REQUIRE(in_dtype == LAFS_F32 || in_scale != 0.f, LAFS_ERR_ARG, "lafs_gather_embed_fwd: in_scale must be non-zero");
  CUtensorMap tw;
  int rc = TmaEncoder::bf16_2d_sw128(&tw, w_perm_bf16, (uint64_t)n_models * dim, pe::kFeat, pe::kFeat * 2, 128);
  if (rc) return rc;
  EmbedParams p{};
  p.imgs = imgs; p.theta = theta; p.bias = bias;
  p.out[0] = out0; p.out[1] = out1;
  LAFS_REQUIRE(((uintptr_t)tokens_perm_out & 15u) == 0, LAFS_ERR_ARG, "lafs_gather_embed_fwd_save: tokens_perm_out misaligned");
  p.tok_out = (__nv_bfloat16*)tokens_perm_out;
  if (pos0 != nullptr) {
    LAFS_REQUIRE(cls0 != nullptr && (n_models == 1 || (pos1 != nullptr && cls1 != nullptr)), LAFS_ERR_ARG,
                 "lafs_gather_embed_seq_fwd: pos / cls of every model are required");
    LAFS_REQUIRE(drop_p >= 0.f && drop_p < 1.f, LAFS_ERR_ARG, "lafs_gather_embed_seq_fwd: drop_p=%f outside [0,1)", drop_p);
    LAFS_REQUIRE((long long)Bv * (n + 1) * dim < (1LL << 32), LAFS_ERR_ARG, "lafs_gather_embed_seq_fwd: output too large for the dropout counter");
    p.seq = 1;
    p.pos[0] = pos0; p.cls[0] = cls0; p.pos[1] = pos1; p.cls[1] = cls1;
    p.drop_p = drop_p; p.drop_scale = 1.f / (1.f - drop_p); p.drop_seed = drop_seed;
  }
  p.Bv = Bv; p.n = n; p.n_pad = (n + 15) & ~15; p.dim = dim; p.n_models = n_models;
  p.mchunks = n_models * dim / 128;
  // faces per group: as many as fit the tile, traded against load balance over the 148 CTAs.
  // cost per CTA ~ (half-)rounds * faces per group (a sweep over G for the 36-landmark views showed the
  // pass over the weights itself is not the limiter: 52-56 us for every G); ties go to the larger
  // group.  The groups of a last, at most half-filled round are split into two half units (first /
  // second half of the weight chunks), which makes that round cost half.
  {
    int gmax = pe::kTokRows / n < 1 ? 1 : pe::kTokRows / n;
    if (gmax > Bv) gmax = Bv;
    long long best_cost = -1;
    int gmin = 1;
    if (const char* e = getenv("LAFS_PE_GFACES")) {   // development override
      const int v = atoi(e);
      if (v >= 1 && v <= gmax) gmin = gmax = v;
    }
    for (int g = gmin; g <= gmax; ++g) {
      const int groups = (Bv + g - 1) / g;
      const int rem = groups % kNumSMs;
      const bool split = groups > kNumSMs && rem > 0 && 2 * rem <= kNumSMs && (p.mchunks % 2 == 0);
      const long long half_rounds = 2LL * (groups / kNumSMs) + (rem == 0 ? 0 : (split ? 1 : 2));
      const long long cost = half_rounds * g;
      if (best_cost < 0 || cost <= best_cost) {
        best_cost = cost;
        p.gfaces = g;
        p.ngroups = groups;
        p.nfull = split ? groups - rem : groups;
        p.nunits = p.nfull + 2 * (groups - p.nfull);
      }
    }
  }
  if (in_dtype == LAFS_U8) { p.in_scale = in_scale; p.in_shift = in_shift; p.pad_raw = -in_shift / in_scale; }
  else { p.in_scale = 1.f; p.in_shift = 0.f; p.pad_raw = 0.f; }
  if (const char* dbg = getenv("LAFS_PE_DEBUG")) p.debug = atoi(dbg);
  cudaStream_t st = (cudaStream_t)stream;
  return in_dtype == LAFS_U8 ? launch_embed<uint8_t>(tw, p, out_dtype, dim, st) : launch_embed<float>(tw, p, out_dtype, dim, st);
}
