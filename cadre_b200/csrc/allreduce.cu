// Gradient exchange of the data-parallel learner: sum of the flat fp32 gradient over the ranks of one NVSwitch
// domain, in place. Replaces Shared_grad_buffers.add_gradient + the chief's "install the summed gradient"
// (ppo_agent/models.py:231-239, ppo_agent/chief.py:13-16) for one process per GPU.
//
// The gradient buffer lives in symmetric memory (same allocation mapped on every rank, plus ONE multicast mapping of
// all of them). Rank r owns the r-th slice of a range: it pulls the slice with multimem.ld_reduce (the switch adds the
// W copies in flight and returns one fp32 sum per element) and pushes the sum back with multimem.st (the switch
// writes all W copies). Every element is reduced exactly once and every rank receives that one result, so replicas
// stay bit-identical by construction; each GPU moves S bytes per direction instead of the ring's 2(W-1)/W * S, and no
// staging buffer is involved. Without a multicast mapping the same kernel sums plain peer loads in rank order and
// stores the result to every peer.
//
// Ordering: block b of every rank pairs with block b of every other rank through flags in symmetric memory
// (flag[b][src] on the destination rank, monotonically increasing epochs, release / acquire at system scope): an entry
// barrier (all ranks have finished producing the range: the kernel is stream-ordered behind the local producer), the
// slice, a system fence, an exit barrier (all slices have landed everywhere). Blocks never wait for other blocks of
// their own grid, so a partially resident grid cannot dead-lock; polling is bounded by a clock budget and a lost
// signal sets an error word instead of hanging the GPU.
#include "../../include/cadre_b200.h"

#include "internal.h"

#include <cstdlib>
#include <memory>

namespace cadre {
namespace {

constexpr int AR_THREADS = 256;
constexpr int AR_UNROLL = 8;
constexpr int AR_MAX_BLOCKS = 256;
constexpr int AR_MAX_WORLD = 16;
constexpr long long AR_POLL_BUDGET = 4000000000ll;   // ~2 s of SM clock

struct ArParams {
  float* local;              // this rank's mapping of the symmetric buffer
  float* multicast;          // multicast mapping (nullptr: peer loads / stores)
  float* peers[AR_MAX_WORLD];
  uint32_t* flags[AR_MAX_WORLD];   // flags[r] = rank r's flag array [AR_MAX_BLOCKS][AR_MAX_WORLD], mapped here
  uint32_t* epoch;           // local [AR_MAX_BLOCKS]
  int* error;                // local
  long long offset, count;   // floats; both multiples of 4
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 ld_peer(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(float* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// threads 0..world-1 of block b: tell rank `t` that this rank reached `value`, then wait for rank `t`
__device__ __forceinline__ void pair_barrier(const ArParams& p, int b, uint32_t value) {
  const int t = threadIdx.x;
  if (t < p.world) {
    st_release_sys(p.flags[t] + b * AR_MAX_WORLD + p.rank, value);
    const uint32_t* mine = p.flags[p.rank] + b * AR_MAX_WORLD + t;
    const long long t0 = clock64();
    while (static_cast<int>(ld_acquire_sys(mine) - value) < 0) {
      __nanosleep(20);
      if (clock64() - t0 > AR_POLL_BUDGET) {
        atomicExch(p.error, 1 + t);
        break;
      }
    }
  }
  __syncthreads();
}

template <bool MC>
__global__ void __launch_bounds__(AR_THREADS) allreduce_kernel(const ArParams p) {
  const int b = blockIdx.x, nb = gridDim.x;
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = p.epoch[b];
  __syncthreads();
  const uint32_t e = s_epoch;
  pair_barrier(p, b, e + 1);

  const long long nv = p.count >> 2;
  const long long per = (nv + p.world - 1) / p.world;
  const long long v0 = per * p.rank, v1 = (v0 + per < nv) ? v0 + per : nv;
  const long long stride = static_cast<long long>(nb) * AR_THREADS;
  auto reduce_one = [&](long long f) -> float4 {
    if (MC) return multimem_ld_reduce_add(p.multicast + f);
    float4 acc = ld_peer(p.peers[0] + f);
    for (int r = 1; r < p.world; ++r) {
      const float4 x = ld_peer(p.peers[r] + f);
      acc.x += x.x, acc.y += x.y, acc.z += x.z, acc.w += x.w;
    }
    return acc;
  };
  auto publish_one = [&](long long f, const float4& v) {
    if (MC) {
      multimem_st(p.multicast + f, v);
    } else {
      for (int r = 0; r < p.world; ++r) st_peer(p.peers[r] + f, v);
    }
  };
  long long i = v0 + static_cast<long long>(b) * AR_THREADS + threadIdx.x;
  // full batches: AR_UNROLL independent 16-byte reductions in flight per thread before the first store
  for (; i + (AR_UNROLL - 1) * stride < v1; i += stride * AR_UNROLL) {
    const long long f = p.offset + (i << 2);
    float4 acc[AR_UNROLL];
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) acc[u] = reduce_one(f + ((u * stride) << 2));
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) publish_one(f + ((u * stride) << 2), acc[u]);
  }
  for (; i < v1; i += stride) {
    const long long f = p.offset + (i << 2);
    publish_one(f, reduce_one(f));
  }
  __threadfence_system();
  __syncthreads();
  pair_barrier(p, b, e + 2);
  if (threadIdx.x == 0) p.epoch[b] = e + 2;
}

struct ArHandle {
  ArParams p;
  int blocks = 32;
  int device = 0;
  long long total = 0;
};

}  // namespace
}  // namespace cadre

#define CADRE_API_BEGIN try {
#define CADRE_API_END                \
  }                                  \
  catch (const cadre::Error& e) {    \
    cadre::set_last_error(e.what()); \
    return e.code;                   \
  }                                  \
  catch (const std::exception& e) {  \
    cadre::set_last_error(e.what()); \
    return 99;                       \
  }                                  \
  return 0;

extern "C" {

int cadre_allreduce_flag_bytes(void) {
  return cadre::AR_MAX_BLOCKS * cadre::AR_MAX_WORLD * static_cast<int>(sizeof(uint32_t));
}

int cadre_allreduce_create(void** handle, int rank, int world, float* const* buffer_ptrs_host, float* multicast,
                           uint32_t* const* flag_ptrs_host, int64_t count) {
  CADRE_API_BEGIN
  using namespace cadre;
  CADRE_REQUIRE(handle != nullptr && buffer_ptrs_host != nullptr && flag_ptrs_host != nullptr, "null argument");
  CADRE_REQUIRE(world >= 2 && world <= AR_MAX_WORLD && rank >= 0 && rank < world, "rank / world");
  CADRE_REQUIRE(count > 0 && count % 4 == 0, "count must be a positive multiple of 4 floats");
  auto h = std::make_unique<ArHandle>();
  h->device = current_device();
  h->total = count;
  ArParams& p = h->p;
  memset(&p, 0, sizeof(p));
  p.rank = rank, p.world = world, p.multicast = multicast;
  for (int r = 0; r < world; ++r) {
    CADRE_REQUIRE(buffer_ptrs_host[r] != nullptr && flag_ptrs_host[r] != nullptr, "peer mapping");
    CADRE_REQUIRE((reinterpret_cast<uintptr_t>(buffer_ptrs_host[r]) & 15) == 0, "16-byte aligned buffers");
    p.peers[r] = buffer_ptrs_host[r], p.flags[r] = flag_ptrs_host[r];
  }
  CADRE_REQUIRE(multicast == nullptr || (reinterpret_cast<uintptr_t>(multicast) & 15) == 0, "16-byte aligned multicast");
  p.local = buffer_ptrs_host[rank];
  CADRE_CUDA_CHECK(cudaMalloc(&p.epoch, AR_MAX_BLOCKS * sizeof(uint32_t)));
  CADRE_CUDA_CHECK(cudaMemset(p.epoch, 0, AR_MAX_BLOCKS * sizeof(uint32_t)));
  CADRE_CUDA_CHECK(cudaMalloc(&p.error, sizeof(int)));
  CADRE_CUDA_CHECK(cudaMemset(p.error, 0, sizeof(int)));
  CADRE_CUDA_CHECK(cudaDeviceSynchronize());
  if (const char* e = getenv("CADRE_AR_BLOCKS")) h->blocks = atoi(e);
  CADRE_REQUIRE(h->blocks >= 1 && h->blocks <= AR_MAX_BLOCKS, "CADRE_AR_BLOCKS out of range");
  *handle = h.release();
  CADRE_API_END
}

int cadre_allreduce_set_blocks(void* handle, int blocks) {
  CADRE_API_BEGIN
  auto* h = static_cast<cadre::ArHandle*>(handle);
  CADRE_REQUIRE(h != nullptr, "handle");
  CADRE_REQUIRE(blocks >= 1 && blocks <= cadre::AR_MAX_BLOCKS, "blocks out of range");
  h->blocks = blocks;
  CADRE_API_END
}

int cadre_allreduce_destroy(void* handle) {
  CADRE_API_BEGIN
  auto* h = static_cast<cadre::ArHandle*>(handle);
  if (h) {
    cudaFree(h->p.epoch);
    cudaFree(h->p.error);
    delete h;
  }
  CADRE_API_END
}

int cadre_allreduce_sum(void* handle, int64_t offset, int64_t count, int use_multicast, void* stream) {
  CADRE_API_BEGIN
  using namespace cadre;
  auto* h = static_cast<ArHandle*>(handle);
  CADRE_REQUIRE(h != nullptr, "handle");
  CADRE_REQUIRE(offset >= 0 && count > 0 && offset % 4 == 0 && count % 4 == 0 && offset + count <= h->total,
                "range must be 16-byte aligned and inside the buffer");
  CADRE_REQUIRE(current_device() == h->device, "all-reduce handle used on another device");
  ArParams p = h->p;
  p.offset = offset, p.count = count;
  auto s = static_cast<cudaStream_t>(stream);
  if (use_multicast) {
    CADRE_REQUIRE(p.multicast != nullptr, "no multicast mapping for this buffer");
    allreduce_kernel<true><<<h->blocks, AR_THREADS, 0, s>>>(p);
  } else {
    allreduce_kernel<false><<<h->blocks, AR_THREADS, 0, s>>>(p);
  }
  CADRE_CUDA_CHECK(cudaGetLastError());
  CADRE_API_END
}

int cadre_allreduce_check(void* handle) {
  CADRE_API_BEGIN
  using namespace cadre;
  auto* h = static_cast<ArHandle*>(handle);
  CADRE_REQUIRE(h != nullptr, "handle");
  CADRE_CUDA_CHECK(cudaDeviceSynchronize());
  int err = 0;
  CADRE_CUDA_CHECK(cudaMemcpy(&err, h->p.error, sizeof(int), cudaMemcpyDeviceToHost));
  if (err != 0)
    throw Error(5, "gradient all-reduce gave up waiting for rank " + std::to_string(err - 1) +
                       " (a peer never reached the barrier)");
  CADRE_API_END
}

}  // extern "C"
