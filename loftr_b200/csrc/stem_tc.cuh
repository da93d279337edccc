// Backbone stem on the tensor cores: conv 7x7 stride 2, 1 -> 128 channels, + folded BatchNorm + ReLU
// (reference src/loftr/backbone/resnet_fpn.py:58-60,101), written as NHWC fp16 hi/lo planes.
//
// As a GEMM the stem is M = output pixels, N = 128, K = 49 (padded to 64): far too little K for the TMA-fed
// gemm_split_kernel, and a single-channel image cannot be im2col'ed by a tensor map (the tap stride would be 2 or 4
// bytes).  So this kernel builds the A tile in software: per 8 x 16 output-pixel tile the 21 x 37 input patch is
// staged in shared memory, every warp writes its rows of the [128 x 64] fp16 hi/lo operand tiles straight into the
// 128-byte-swizzled K-major layout tcgen05.mma expects, and one thread issues the 12 MMAs (4 k-steps x (hi*hi + hi*lo +
// lo*hi)) into one of two TMEM stages while all 8 warps run the epilogue of the previous tile (BN + ReLU + split,
// TMA store of 2 x 16 pixel x 32 channel boxes).  The weights (128 x 49 fp32) are scaled by a power of two, split and
// swizzled into shared memory once per CTA.  One persistent CTA per SM, 256 threads.
#pragma once
#include "epilogues.cuh"

namespace lb {

struct StemTcParams {
  const float* img;     // [N, 1, H, W]
  int N, H, W;
  const float* wt;      // conv1.weight transposed [49][128]
  const float* scale;   // folded bn1 [128]
  const float* shift;
  OutMaps om;           // NHWC planes [N, H/2, W/2, 128]: 4-D TMA-store maps (box 32 ch x 16 x 2 px)
  int tiles_w, tiles_h; // 16- / 8-pixel tiles of the output grid
};

constexpr int kStemThreads = 256;
constexpr int kStemPatchH = 2 * kConvTileH + 5;    // 21 input rows feed 8 output rows
constexpr int kStemPatchW = 2 * kConvTileW + 5;    // 37
constexpr int kStemPatchPitch = 40;
constexpr int kStemATile = 128 * 128;              // [128 rows][64 fp16] = 16 KB per plane
constexpr int kStemSmemBytes = 4 * kStemATile      // A hi/lo x 2 buffers
                               + 2 * kStemATile    // W hi/lo
                               + 8 * 4096          // per-warp TMA-store staging
                               + 2 * kStemPatchH * kStemPatchPitch * 4 + 2 * 128 * 4 + 64 * 4 + 64;

__device__ __forceinline__ uint32_t stem_sw128(int row, int k) {   // byte offset of fp16 element (row, k) in a SW128 K-major tile
  return static_cast<uint32_t>(row * 128 + ((((k >> 3) ^ (row & 7)) << 4) | ((k & 7) << 1)));
}

__global__ void __launch_bounds__(kStemThreads, 1) conv_stem7x7_tc_kernel(const __grid_constant__ StemTcParams p) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the first convolution may set up meanwhile
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) asm volatile("trap;");
  uint8_t* sA = smem;                                   // [buf][hi|lo][16 KB]
  uint8_t* sW = smem + 4 * kStemATile;                  // [hi|lo][16 KB]
  uint32_t* sStage = reinterpret_cast<uint32_t*>(smem + 6 * kStemATile);
  float* sPatch = reinterpret_cast<float*>(smem + 6 * kStemATile + 8 * 4096);   // [2][21][40]
  float* sScale = sPatch + 2 * kStemPatchH * kStemPatchPitch;
  float* sShift = sScale + 128;
  float* sRed = sShift + 128;                           // [64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 64);                      // [2] accumulator-ready barriers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 256);

  // ---- weights: power-of-two scale (keeps the fp16 `lo` residuals out of the subnormal range), split, swizzle
  float amax = 0.f;
  for (int i = tid; i < 49 * 128; i += kStemThreads) amax = fmaxf(amax, fabsf(p.wt[i]));
  for (int o = 16; o; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  if (lane == 0) sRed[warp] = amax;
  __syncthreads();
  amax = sRed[0];
  for (int k = 1; k < kStemThreads / 32; ++k) amax = fmaxf(amax, sRed[k]);
  int e = amax > 0.f ? static_cast<int>(floorf(log2f(4096.f / amax))) : 0;
  e = max(-24, min(24, e));
  const float wmul = exp2f(static_cast<float>(e)), inv = exp2f(static_cast<float>(-e));
  for (int i = tid; i < 128 * 32; i += kStemThreads) {          // (cout n, k pair)
    const int n = i >> 5, k = (i & 31) << 1;
    const float w0 = k < 49 ? p.wt[k * 128 + n] * wmul : 0.f;
    const float w1 = k + 1 < 49 ? p.wt[(k + 1) * 128 + n] * wmul : 0.f;
    uint32_t h, l;
    split_f16x2(w0, w1, h, l);
    *reinterpret_cast<uint32_t*>(sW + stem_sw128(n, k)) = h;
    *reinterpret_cast<uint32_t*>(sW + kStemATile + stem_sw128(n, k)) = l;
  }
  for (int i = tid; i < 4 * kStemATile / 16; i += kStemThreads) reinterpret_cast<uint4*>(sA)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < 128) {
    sScale[tid] = p.scale[tid] * inv;
    sShift[tid] = p.shift[tid];
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int Ho = p.H / 2, Wo = p.W / 2;
  const long tiles_per_img = static_cast<long>(p.tiles_w) * p.tiles_h;
  const long total = tiles_per_img * p.N;
  constexpr uint32_t idesc = umma_idesc_f16_f32(128, 128);
  const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
  const int half = warp >> 2;                     // column half of the accumulator this warp drains
  uint32_t* stage = sStage + warp * 1024;

  uint32_t phase[2] = {0u, 0u};
  long prev = -1;
  int it = 0;
  for (long t = blockIdx.x;; t += gridDim.x, ++it) {
    const bool have = t < total;
    const int b = it & 1;
    if (have) {
      const int n = static_cast<int>(t / tiles_per_img);
      const int rem = static_cast<int>(t - n * tiles_per_img);
      const int ty = rem / p.tiles_w, tx = rem - ty * p.tiles_w;
      // (a) input patch -> shared memory (zero outside the image = the convolution's padding)
      float* patch = sPatch + b * kStemPatchH * kStemPatchPitch;
      const int iy0 = 2 * ty * kConvTileH - 3, ix0 = 2 * tx * kConvTileW - 3;
      for (int i = tid; i < kStemPatchH * kStemPatchW; i += kStemThreads) {
        const int py = i / kStemPatchW, px = i - py * kStemPatchW;
        const int iy = iy0 + py, ix = ix0 + px;
        float v = 0.f;
        if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) v = p.img[(static_cast<long>(n) * p.H + iy) * p.W + ix];
        patch[py * kStemPatchPitch + px] = v;
      }
      __syncthreads();
      // (b) software im2col: warp w writes rows w, w+8, ...; lane = pair of taps (k, k+1)
      uint8_t* a_hi = sA + b * 2 * kStemATile;
      uint8_t* a_lo = a_hi + kStemATile;
      if (lane < 25) {
        const int k0 = 2 * lane, k1 = k0 + 1;
        const int ky0 = k0 / 7, kx0 = k0 - ky0 * 7;
        const int ky1 = k1 / 7, kx1 = k1 - ky1 * 7;
        for (int r = warp; r < 128; r += kStemThreads / 32) {
          const int py = r >> 4, px = r & 15;
          const float v0 = patch[(2 * py + ky0) * kStemPatchPitch + 2 * px + kx0];
          const float v1 = k1 < 49 ? patch[(2 * py + ky1) * kStemPatchPitch + 2 * px + kx1] : 0.f;
          uint32_t h, l;
          split_f16x2(v0, v1, h, l);
          *reinterpret_cast<uint32_t*>(a_hi + stem_sw128(r, k0)) = h;
          *reinterpret_cast<uint32_t*>(a_lo + stem_sw128(r, k0)) = l;
        }
      }
      fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();          // A(b) complete; every warp has finished draining the TMEM stage that MMA(b) overwrites
    if (have && tid == 0) {
      tc_fence_after();
      const uint32_t a0 = smem_u32(sA + b * 2 * kStemATile), w0 = smem_u32(sW);
      const uint64_t da_hi = umma_desc_k_sw128(a0), da_lo = umma_desc_k_sw128(a0 + kStemATile);
      const uint64_t db_hi = umma_desc_k_sw128(w0), db_lo = umma_desc_k_sw128(w0 + kStemATile);
      const uint32_t d = tmem_base + b * 128;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t adv = static_cast<uint64_t>(k * 2);
        umma_f16(d, da_hi + adv, db_hi + adv, idesc, k != 0 ? 1u : 0u);
        umma_f16(d, da_hi + adv, db_lo + adv, idesc, 1u);
        umma_f16(d, da_lo + adv, db_hi + adv, idesc, 1u);
      }
      umma_commit(&bars[b]);
    }
    // (c) epilogue of the previous tile while the tensor core works on this one
    if (prev >= 0) {
      const int pb = b ^ 1;
      mbar_wait(&bars[pb], phase[pb]);
      phase[pb] ^= 1u;
      tc_fence_after();
      const int n = static_cast<int>(prev / tiles_per_img);
      const int rem = static_cast<int>(prev - n * tiles_per_img);
      const int ty = rem / p.tiles_w, tx = rem - ty * p.tiles_w;
#pragma unroll 1
      for (int c = half * 2; c < half * 2 + 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tmem_base + pb * 128 + c * 32 + lane_base, v);
        tmem_ld_wait();
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = fmaxf(fmaf(__uint_as_float(v[j]), sScale[c * 32 + j], sShift[c * 32 + j]), 0.f);
        warp_tma_store_planes32(stage, p.om, OutCoord{c * 32, tx * kConvTileW, ty * kConvTileH + (warp & 3) * 2, n}, x);
      }
    }
    prev = have ? t : -1;
    if (!have) break;
  }
  (void)Ho;
  (void)Wo;
  tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace lb
