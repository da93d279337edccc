// Multi-GPU exchange of the match lists: pairs are sharded over ranks, every rank ends with the global list.
// Replaces the reference's `gather()` (src/utils/comm.py:113-176: size exchange + padded pickled-object all_gather on
// a gloo side group, called from src/lightning/lightning_loftr.py:235,241) with ONE static-shape ncclAllGather of a
// packed float32 wire buffer on the compute stream, between a pack and an unpack kernel.
//
// Wire format per rank: [1 + capacity][6] float32
//     row 0     : (count, 0, 0, 0, 0, 0)
//     row 1 + k : (x0, y0, x1, y1, mconf, global pair id)      k < count
// NCCL is bound at run time (dlopen of libnccl.so.2 -- the copy the host process already has loaded, e.g. PyTorch's)
// so the library itself has no link-time dependency on it.
#pragma once
#include <dlfcn.h>

namespace lb {

constexpr int kWireCols = 6;

__global__ void pack_matches_kernel(const float* __restrict__ mk0, const float* __restrict__ mk1,
                                    const float* __restrict__ conf, const long long* __restrict__ bids, long count,
                                    int pair_offset, float* __restrict__ wire, long capacity) {
  const long k = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (k == 0) {
    wire[0] = static_cast<float>(count);
    for (int c = 1; c < kWireCols; ++c) wire[c] = 0.f;
  }
  if (k >= count || k >= capacity) return;
  float* row = wire + (1 + k) * kWireCols;
  row[0] = mk0[2 * k];
  row[1] = mk0[2 * k + 1];
  row[2] = mk1[2 * k];
  row[3] = mk1[2 * k + 1];
  row[4] = conf[k];
  row[5] = static_cast<float>(bids[k] + pair_offset);
}

// gathered [world][1 + capacity][6] -> concatenated lists in rank order (ranks hold contiguous pair blocks, so this is
// the reference's ascending (pair, i) order).  counts_out: [world + 1] = per-rank counts and the total.
__global__ void unpack_matches_kernel(const float* __restrict__ gathered, int world, long capacity,
                                      float* __restrict__ mk0, float* __restrict__ mk1, float* __restrict__ conf,
                                      long long* __restrict__ bids, long out_capacity, int* __restrict__ counts_out) {
  const int r = blockIdx.y;
  const long stride = (1 + capacity) * kWireCols;
  long offset = 0, total = 0;
  long mine = 0;
  for (int q = 0; q < world; ++q) {
    long c = static_cast<long>(gathered[q * stride] + 0.5f);
    if (c > capacity) c = capacity;   // overflow is reported through counts_out (the true count) and handled by the host
    if (q < r) offset += c;
    if (q == r) mine = c;
    total += c;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    counts_out[r] = static_cast<int>(gathered[r * stride] + 0.5f);
    if (r == 0) counts_out[world] = static_cast<int>(total);
  }
  const long k = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (k >= mine) return;
  const long o = offset + k;
  if (o >= out_capacity) return;
  const float* row = gathered + r * stride + (1 + k) * kWireCols;
  mk0[2 * o] = row[0];
  mk0[2 * o + 1] = row[1];
  mk1[2 * o] = row[2];
  mk1[2 * o + 1] = row[3];
  conf[o] = row[4];
  bids[o] = static_cast<long long>(row[5] + 0.5f);
}

// ------------------------------------------------------------------------------------------------ NCCL binding
struct NcclApi {
  typedef struct { char internal[128]; } UniqueId;
  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  void* handle = nullptr;
};
constexpr int kNcclFloat32 = 7;   // ncclFloat32 in nccl.h's ncclDataType_t

static NcclApi* nccl_api(const char* path_hint) {
  static NcclApi api;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (api.handle) return &api;
  const char* env = getenv("LOFTR_B200_NCCL_LIB");
  const char* candidates[] = {env, path_hint, "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* c : candidates) {
    if (!c || !*c) continue;
    h = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) return nullptr;
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
  if (!api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy) {
    dlclose(h);
    return nullptr;
  }
  api.handle = h;
  return &api;
}

struct LbComm {
  void* nccl;
  int rank, world, device;
};

}  // namespace lb
