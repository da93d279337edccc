// KV = K^T V of the coarse linear attention on the tensor cores (reference linear_attention.py:43:
// `KV = einsum("nshd,nshv->nhdv", K, values)`), fed from the fp16 hi/lo planes the k|v projection epilogue
// (EpiKvProj) writes.
//
// As a GEMM this is D[d, v] = sum_rows K[row, d] * V[row, v]: the contraction runs over the ROWS of two row-major
// matrices, i.e. both operands are "MN-major" for tcgen05 (idesc a_major = b_major = 1).  A TMA box of 64 columns x
// 64 rows of the planes lands in shared memory as [64 k-rows][128 B] with the 128-byte swizzle, which is exactly the
// canonical MN-major SWIZZLE_128B layout  ((8,n),(8,k)) : ((1,LBO),(8,SBO))  in 16-byte units (CuTe make_umma_desc<MN>):
// SBO = 1024 B between groups of 8 k-rows, LBO = 8192 B between blocks of 64 d (or v), 2048 B per K = 16 MMA step.
//
// One CTA per (k-split, head quartet, image): M = 128 (the d of 4 heads), N = 128 (the v of the same 4 heads),
// K = its share of the image's rows (rows past the image end are zero-filled by TMA, masked rows are zero in the
// planes).  Only the four diagonal 32 x 32 blocks of the 128 x 128 product are attention state; the epilogue stores
// those as one partial per (image, split, head).  Split precision as everywhere: hi*hi into the main accumulator,
// hi*lo + lo*hi into the correction accumulator, summed once.
//
// Ksum = K^T 1 (`K.sum(dim=1)`, linear_attention.py:44) rides along: two N = 16 MMAs per K step multiply the K planes
// with a block of ones (1 KB of fp16 1.0 in shared memory -- with every element equal, any descriptor that stays inside
// the block is a valid "ones" operand) into 16 more accumulator columns; masked / out-of-range rows are already zero
// in the planes.
#pragma once
#include "ptx.cuh"

namespace lb {

struct KvGemmParams {
  float* part;        // [images][splits][H = 8][32*32 + 32]: KV then Ksum, the layout of the final kv state
  int kb_total;       // ceil(rows per image / 64)
  int kb_per_split;
  int splits;
};

constexpr int kKvGemmThreads = 192;                 // warp 0 TMA, warp 1 MMA + TMEM, warps 2-5 epilogue
constexpr int kKvGemmStages = 3;
constexpr int kKvBlock = 64 * 128;                  // one TMA box: 64 k-rows x 128 bytes
constexpr int kKvTile = 2 * kKvBlock;               // 128 d (or v) x 64 k-rows per plane: 16 KB
constexpr int kKvStageBytes = 4 * kKvTile;          // A hi, A lo, B hi, B lo
constexpr int kKvOnesBytes = 1024;
constexpr int kKvGemmSmem = kKvGemmStages * kKvStageBytes + kKvOnesBytes + 256;

// MN-major operand tile, 128-byte swizzle: LBO = 8192 B (next 64 MN elements), SBO = 1024 B (next 8 k-rows)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(kKvBlock >> 4) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// kind::f16 instruction descriptor with both operands MN-major (bits 15, 16)
__host__ __device__ constexpr uint32_t umma_idesc_f16_f32_mn(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// A MN-major, B K-major (the ones block)
__host__ __device__ constexpr uint32_t umma_idesc_f16_f32_mn_k(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 15) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
// unswizzled K-major descriptor into the ones block: 8-row groups 128 B apart, the two 16-byte K halves 256 B apart
__device__ __forceinline__ uint64_t umma_desc_ones(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(256 >> 4) << 16;
  d |= static_cast<uint64_t>(128 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  return d;
}

// tm_hi / tm_lo: the K|V planes [images][rows][512] with box (64 columns, 64 rows, 1); columns as the projection emits
// them: [K of heads 0-3 | V of heads 0-3 | K of heads 4-7 | V of heads 4-7]
__global__ void __launch_bounds__(kKvGemmThreads, 1)
kv_gemm_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, const KvGemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) asm volatile("trap;");
  uint8_t* s_ones = smem + kKvGemmStages * kKvStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_ones + kKvOnesBytes);
  uint64_t* empty_bar = full_bar + kKvGemmStages;
  uint64_t* acc_bar = empty_bar + kKvGemmStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int split = blockIdx.x, quartet = blockIdx.y, image = blockIdx.z;
  const int kb0 = split * p.kb_per_split;
  const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
  const int nkb = max(kb1 - kb0, 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_hi);
    tma_prefetch_desc(&tm_lo);
    for (int s = 0; s < kKvGemmStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  for (int i = threadIdx.x; i < kKvOnesBytes / 4; i += kKvGemmThreads) reinterpret_cast<uint32_t*>(s_ones)[i] = 0x3C003C00u;   // fp16 1.0 x 2
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + stage * kKvStageBytes;
        uint64_t* fb = &full_bar[stage];
        mbar_arrive_expect_tx(fb, kKvStageBytes);
        const int row0 = (kb0 + i) * 64;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int ck = quartet * 256 + b * 64, cv = quartet * 256 + 128 + b * 64;
          tma_load_3d(st + b * kKvBlock, &tm_hi, fb, ck, row0, image);
          tma_load_3d(st + kKvTile + b * kKvBlock, &tm_lo, fb, ck, row0, image);
          tma_load_3d(st + 2 * kKvTile + b * kKvBlock, &tm_hi, fb, cv, row0, image);
          tma_load_3d(st + 3 * kKvTile + b * kKvBlock, &tm_lo, fb, cv, row0, image);
        }
        if (++stage == kKvGemmStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16_f32_mn(128, 128);
      constexpr uint32_t idesc_ones = umma_idesc_f16_f32_mn_k(128, 16);
      const uint32_t d_main = tmem_base, d_corr = tmem_base + 128, d_ksum = tmem_base + 256;
      const uint64_t ones = umma_desc_ones(smem_u32(s_ones));
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < nkb; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + stage * kKvStageBytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t off = static_cast<uint32_t>(k) * 2048u;   // 16 k-rows of 128 bytes
          const uint64_t a_hi = umma_desc_mn_sw128(st + off), a_lo = umma_desc_mn_sw128(st + kKvTile + off);
          const uint64_t b_hi = umma_desc_mn_sw128(st + 2 * kKvTile + off), b_lo = umma_desc_mn_sw128(st + 3 * kKvTile + off);
          const uint32_t not_first = (i | k) != 0 ? 1u : 0u;
          umma_f16(d_main, a_hi, b_hi, idesc, not_first);
          umma_f16(d_corr, a_hi, b_lo, idesc, not_first);
          umma_f16(d_corr, a_lo, b_hi, idesc, 1u);
          umma_f16(d_ksum, a_hi, ones, idesc_ones, not_first);
          umma_f16(d_ksum, a_lo, ones, idesc_ones, 1u);
        }
        umma_commit(&empty_bar[stage]);
        if (++stage == kKvGemmStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      umma_commit(acc_bar);
    }
  } else {
    // ---- epilogue: TMEM lane = d of the quartet (lane quarter = head), this thread keeps its head's 32 v columns
    const int q = warp & 3;                     // TMEM lane quarter this warp may read = head within the quartet
    const int d = lane;
    float* base = p.part + ((static_cast<long>(image) * p.splits + split) * 8 + quartet * 4 + q) * 1056;
    float* out = base + d * 32;
    if (nkb > 0) {
      mbar_wait(acc_bar, 0);
      tc_fence_after();
      uint32_t v[32], w[32], ks[32];
      const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
      tmem_ld32(tmem_base + q * 32 + lane_base, v);
      tmem_ld32(tmem_base + 128 + q * 32 + lane_base, w);
      tmem_ld32(tmem_base + 256 + lane_base, ks);     // 16 identical columns of Ksum (+ 16 unused ones)
      tmem_ld_wait();
      base[1024 + d] = __uint_as_float(ks[0]);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<float4*>(out)[j] =
            make_float4(__uint_as_float(v[4 * j]) + __uint_as_float(w[4 * j]), __uint_as_float(v[4 * j + 1]) + __uint_as_float(w[4 * j + 1]),
                        __uint_as_float(v[4 * j + 2]) + __uint_as_float(w[4 * j + 2]), __uint_as_float(v[4 * j + 3]) + __uint_as_float(w[4 * j + 3]));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) reinterpret_cast<float4*>(out)[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      base[1024 + d] = 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace lb
