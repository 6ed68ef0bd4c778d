// Data-parallel gradient exchange as ONE hand-written kernel over NVSwitch multicast (NVLS) memory.
//
// The flat fp32 gradient buffer of every rank lives in CUDA symmetric memory that is also mapped through a multicast
// address (allocated and rendezvous'd by the host layer, tf_kaldi_speaker_b200/parallel.py).  Rank r owns the r-th
// slice of the buffer: `multimem.ld_reduce.add.v4.f32` returns the sum over all ranks of 16 bytes, reduced inside the
// switch, and `multimem.st.v4.f32` writes the result back to every rank's copy -- a two-shot all-reduce in which each
// gradient byte crosses each GPU's NVLink once per direction, with no ring / tree pipeline, no staging buffers and no
// host involvement: the kernel sits inside the captured CUDA graph of the step, between the backward pass and the
// optimizer.  Ranks synchronise through per-block flags in a second small symmetric buffer (system-scope
// release / acquire), so every rank must launch this kernel the same number of times with the same grid.
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "xv_internal.h"

#ifndef XV_AR_P2P_CFG_DEFAULT
#define XV_AR_P2P_CFG_DEFAULT 1      // (1024 threads, 4 x 16 B in flight per peer): 77 us for 39 MB at N = 2; the other
                                     // configurations (512x4, 512x8, 1024x2) measured 79-80 us (29 us instead of 45 us at 2 MB)
#endif
#ifndef XV_AR_CFG_DEFAULT
#define XV_AR_CFG_DEFAULT 4          // 256 threads, one 16-byte vector per step, next ld_reduce issued before the current store:
                                     // 8 x B200, 39 MB: 106-107 us against 131 us for "whole slice in flight at once" (cfg 0-3) --
                                     // profiles/r02_allreduce_bench_8gpu.json; anything with <= 1.2 MB in flight per GPU lands there
#endif
#ifndef XV_DP_BARRIER_TIMEOUT_S
#define XV_DP_BARRIER_TIMEOUT_S 20ull      // rank skew at the exchange is micro- to milliseconds; first-step JIT / capture < 20 s
#endif

namespace xv {

// XV_AR_CFG = 0..3: (512,4) / (1024,4) / (512,8) / (1024,8) threads x 16-byte accesses per thread, every load of a thread
// issued before its first store (round 1); 4..12: the software-pipelined kernel below with (256,1) (256,2) (512,1) (512,2)
// (1024,1) (128,2) (128,1) (64,1) (64,2); 13: the round-1 kernel with (256,1).  Measured on 8 x B200: the fewer bytes a GPU
// keeps in flight (down to ~0.6 MB) the better -- with the whole slice issued at once every rank first saturates its
// OUTBOUND link (serving the other ranks' reductions) and then its INBOUND link (receiving the broadcasts); short
// dependent load -> store chains interleave the two phases across warps and keep both directions busy.

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
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// Block-level barrier across all ranks: block b of rank r writes `epoch` into slot [b][r] of every peer's flag buffer
// and waits until its own slots [b][*] have reached `epoch` (flags only grow, so nothing is ever reset).
__device__ __forceinline__ void rank_barrier(uint32_t* const* flags, int rank, int world, uint32_t epoch) {
  __syncthreads();
  if (threadIdx.x < world) {
    const int peer = threadIdx.x;
    __threadfence_system();
    st_release_sys(flags[peer] + blockIdx.x * world + rank, epoch);
    const uint32_t* mine = flags[rank] + blockIdx.x * world + peer;
    // Bounded wait: a rank that died (or never launched this kernel) must not hang the other seven GPUs forever -- after
    // XV_DP_BARRIER_TIMEOUT_S seconds the waiting block traps, which surfaces as a sticky CUDA error on this rank
    // (the same policy as the GEMM's mbarrier waits, xv_ptx.cuh).
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while (ld_acquire_sys(mine) < epoch) {
      if ((++spins & 0xfff) == 0) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > XV_DP_BARRIER_TIMEOUT_S * 1000000000ull) {
          printf("xv: rank barrier timeout (rank %d waits for rank %d, block %d, epoch %u)\n", rank, peer, blockIdx.x, epoch);
          __trap();
        }
      }
    }
  }
  __syncthreads();
}

template <int AR_THREADS, int AR_UNROLL>
__global__ void __launch_bounds__(AR_THREADS) dp_allreduce_multimem_kernel(float* __restrict__ mc, uint32_t* const* __restrict__ flags,
                                                                           uint32_t* __restrict__ block_epoch, int rank, int world,
                                                                           long long n_vec) {
  // per-block launch counter (local memory): the same on every rank because every rank launches the same sequence
  const uint32_t epoch0 = block_epoch[blockIdx.x];
  rank_barrier(flags, rank, world, epoch0 + 1);          // every rank has finished writing its gradients
  const long long per = (n_vec + world - 1) / world;
  const long long lo = per * rank, hi = (lo + per < n_vec) ? lo + per : n_vec;
  const long long stride = static_cast<long long>(gridDim.x) * AR_THREADS;
  for (long long i = lo + static_cast<long long>(blockIdx.x) * AR_THREADS + threadIdx.x; i < hi; i += stride * AR_UNROLL) {
    float4 v[AR_UNROLL];
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u)
      if (i + u * stride < hi) v[u] = multimem_ld_reduce_add(mc + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u)
      if (i + u * stride < hi) multimem_st(mc + 4 * (i + u * stride), v[u]);
  }
  rank_barrier(flags, rank, world, epoch0 + 2);          // every slice has been written back everywhere
  if (threadIdx.x == 0) block_epoch[blockIdx.x] = epoch0 + 2;
}


// Software-pipelined form: a thread walks its vectors U at a time and issues the NEXT group's ld_reduce before it stores the
// current group, so that at any moment part of the grid is pulling reductions (outbound-heavy: every rank serves S bytes)
// while another part is broadcasting results (inbound-heavy) -- both NVLink directions stay busy instead of alternating.
template <int AR_THREADS, int U>
__global__ void __launch_bounds__(AR_THREADS) dp_allreduce_multimem_pipe_kernel(float* __restrict__ mc, uint32_t* const* __restrict__ flags,
                                                                                uint32_t* __restrict__ block_epoch, int rank, int world,
                                                                                long long n_vec) {
  const uint32_t epoch0 = block_epoch[blockIdx.x];
  rank_barrier(flags, rank, world, epoch0 + 1);
  const long long per = (n_vec + world - 1) / world;
  const long long lo = per * rank, hi = (lo + per < n_vec) ? lo + per : n_vec;
  const long long stride = static_cast<long long>(gridDim.x) * AR_THREADS;
  long long i = lo + static_cast<long long>(blockIdx.x) * AR_THREADS + threadIdx.x;
  float4 cur[U], nxt[U];
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (i + u * stride < hi) cur[u] = multimem_ld_reduce_add(mc + 4 * (i + u * stride));
  for (; i < hi; i += stride * U) {
    const long long in = i + stride * U;
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (in + u * stride < hi) nxt[u] = multimem_ld_reduce_add(mc + 4 * (in + u * stride));
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * stride < hi) multimem_st(mc + 4 * (i + u * stride), cur[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) cur[u] = nxt[u];
  }
  rank_barrier(flags, rank, world, epoch0 + 2);
  if (threadIdx.x == 0) block_epoch[blockIdx.x] = epoch0 + 2;
}

// Peer-memory variant (no multicast): rank r sums slice r over the peers' buffers with direct NVLink loads and stores the
// result into every peer's buffer.  Per GPU and direction it moves (N-1)/N of the buffer in each phase -- the same as a
// ring -- so it only wins where the multicast path is wasteful: at N = 2 the in-switch reduction sends a rank's OWN data
// through the switch and back (58 MB per direction instead of 19.5 MB for the 39 MB buffer: 126 us vs 93 us for NCCL).
__device__ __forceinline__ float4 ld_sys_v4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_sys_v4(float* p, float4 v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

template <int AR_THREADS, int AR_UNROLL>
__global__ void __launch_bounds__(AR_THREADS) dp_allreduce_p2p_kernel(float* const* __restrict__ bufs, uint32_t* const* __restrict__ flags,
                                                                      uint32_t* __restrict__ block_epoch, int rank, int world,
                                                                      long long off_vec, long long n_vec) {
  const uint32_t epoch0 = block_epoch[blockIdx.x];
  rank_barrier(flags, rank, world, epoch0 + 1);
  const long long per = (n_vec + world - 1) / world;
  const long long lo = off_vec + per * rank, hi = off_vec + ((per * rank + per < n_vec) ? per * rank + per : n_vec);
  const long long stride = static_cast<long long>(gridDim.x) * AR_THREADS;
  for (long long i = lo + static_cast<long long>(blockIdx.x) * AR_THREADS + threadIdx.x; i < hi; i += stride * AR_UNROLL) {
    float4 acc[AR_UNROLL];
#pragma unroll
    for (int u = 0; u < AR_UNROLL; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < world; ++p) {            // rank order: every slice is summed in the same order on its owner
      const float* src = bufs[p];
      float4 v[AR_UNROLL];
#pragma unroll
      for (int u = 0; u < AR_UNROLL; ++u)
        if (i + u * stride < hi) v[u] = ld_sys_v4(src + 4 * (i + u * stride));
#pragma unroll
      for (int u = 0; u < AR_UNROLL; ++u)
        if (i + u * stride < hi) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
    }
    for (int p = 0; p < world; ++p) {
      float* dst = bufs[p];
#pragma unroll
      for (int u = 0; u < AR_UNROLL; ++u)
        if (i + u * stride < hi) st_sys_v4(dst + 4 * (i + u * stride), acc[u]);
    }
  }
  rank_barrier(flags, rank, world, epoch0 + 2);
  if (threadIdx.x == 0) block_epoch[blockIdx.x] = epoch0 + 2;
}

}  // namespace xv

using namespace xv;

extern "C" int xv_dp_allreduce_multimem(void* multicast_ptr, void* const* flag_ptrs_dev, void* block_epoch, int rank, int world,
                                        int64_t offset, int64_t n, int grid, void* stream) {
  if (!multicast_ptr || !flag_ptrs_dev || !block_epoch || world < 1 || world > 32 || rank < 0 || rank >= world || n <= 0 || (n & 3) ||
      offset < 0 || (offset & 3) || grid < 1 || grid > 1024 || (reinterpret_cast<uintptr_t>(multicast_ptr) & 15))
    return set_error(XV_ERR_INVALID, "xv_dp_allreduce_multimem: bad arguments (n %% 4 == 0, 16-byte aligned multicast pointer, "
                                     "world <= 32, grid <= 1024)");
  int sms; int rc = device_sm_count(&sms); if (rc) return rc;
  if (grid > sms) return set_error(XV_ERR_INVALID, "xv_dp_allreduce_multimem: the grid must be co-resident (grid <= SM count)");
  int cfg = XV_AR_CFG_DEFAULT;      // read per call (a launch is captured once per graph): tools/allreduce_bench.py sweeps it
  if (const char* e = getenv("XV_AR_CFG")) {
    cfg = atoi(e);
    if (cfg < 0 || cfg > 13) cfg = XV_AR_CFG_DEFAULT;
  }
  float* mc = static_cast<float*>(multicast_ptr) + offset;
  uint32_t* const* fl = reinterpret_cast<uint32_t* const*>(flag_ptrs_dev);
  uint32_t* ep = static_cast<uint32_t*>(block_epoch);
  const long long nv = static_cast<long long>(n / 4);
  cudaStream_t s_ = static_cast<cudaStream_t>(stream);
  switch (cfg) {
    case 0: dp_allreduce_multimem_kernel<512, 4><<<grid, 512, 0, s_>>>(mc, fl, ep, rank, world, nv); break;
    case 1: dp_allreduce_multimem_kernel<1024, 4><<<grid, 1024, 0, s_>>>(mc, fl, ep, rank, world, nv); break;
    case 2: dp_allreduce_multimem_kernel<512, 8><<<grid, 512, 0, s_>>>(mc, fl, ep, rank, world, nv); break;
    case 4: dp_allreduce_multimem_pipe_kernel<256, 1><<<grid, 256, 0, s_>>>(mc, fl, ep, rank, world, nv); break;
    case 5: dp_allreduce_multimem_pipe_kernel<256, 2><<<grid, 256, 0, s_>>>(mc, fl, ep, rank, world, nv); break;
    case 6: dp_allreduce_multimem_pipe_kernel<512, 1><<<grid, 512, 0, s_>>>(mc, fl, ep, rank, world, nv); break;
    case 7: dp_allreduce_multimem_pipe_kernel<512, 2><<<grid, 512, 0, s_>>>(mc, fl, ep, rank, world, nv); break;
    case 8: dp_allreduce_multimem_pipe_kernel<1024, 1><<<grid, 1024, 0, s_>>>(mc, fl, ep, rank, world, nv); break;
    case 9: dp_allreduce_multimem_pipe_kernel<128, 2><<<grid, 128, 0, s_>>>(mc, fl, ep, rank, world, nv); break;
    case 10: dp_allreduce_multimem_pipe_kernel<128, 1><<<grid, 128, 0, s_>>>(mc, fl, ep, rank, world, nv); break;
    case 11: dp_allreduce_multimem_pipe_kernel<64, 1><<<grid, 64, 0, s_>>>(mc, fl, ep, rank, world, nv); break;
    case 12: dp_allreduce_multimem_pipe_kernel<64, 2><<<grid, 64, 0, s_>>>(mc, fl, ep, rank, world, nv); break;
    case 13: dp_allreduce_multimem_kernel<256, 1><<<grid, 256, 0, s_>>>(mc, fl, ep, rank, world, nv); break;
    default: dp_allreduce_multimem_kernel<1024, 8><<<grid, 1024, 0, s_>>>(mc, fl, ep, rank, world, nv); break;
  }
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}

extern "C" int xv_dp_allreduce_p2p(void* const* buf_ptrs_dev, void* const* flag_ptrs_dev, void* block_epoch, int rank, int world,
                                   int64_t offset, int64_t n, int grid, void* stream) {
  if (!buf_ptrs_dev || !flag_ptrs_dev || !block_epoch || world < 1 || world > 32 || rank < 0 || rank >= world || n <= 0 || (n & 3) ||
      offset < 0 || (offset & 3) || grid < 1 || grid > 1024)
    return set_error(XV_ERR_INVALID, "xv_dp_allreduce_p2p: bad arguments (n %% 4 == 0, world <= 32, grid <= 1024)");
  int sms; int rc = device_sm_count(&sms); if (rc) return rc;
  if (grid > sms) return set_error(XV_ERR_INVALID, "xv_dp_allreduce_p2p: the grid must be co-resident (grid <= SM count)");
  float* const* bufs = reinterpret_cast<float* const*>(buf_ptrs_dev);
  uint32_t* const* fl = reinterpret_cast<uint32_t* const*>(flag_ptrs_dev);
  static int cfg = -1;
  if (cfg < 0) {
    const char* e = getenv("XV_AR_P2P_CFG");
    cfg = e ? atoi(e) : XV_AR_P2P_CFG_DEFAULT;
    if (cfg < 0 || cfg > 3) cfg = XV_AR_P2P_CFG_DEFAULT;
  }
  uint32_t* ep = static_cast<uint32_t*>(block_epoch);
  const long long ov = static_cast<long long>(offset / 4), nv = static_cast<long long>(n / 4);
  cudaStream_t s_ = static_cast<cudaStream_t>(stream);
  switch (cfg) {
    case 0: dp_allreduce_p2p_kernel<512, 4><<<grid, 512, 0, s_>>>(bufs, fl, ep, rank, world, ov, nv); break;
    case 1: dp_allreduce_p2p_kernel<1024, 4><<<grid, 1024, 0, s_>>>(bufs, fl, ep, rank, world, ov, nv); break;
    case 2: dp_allreduce_p2p_kernel<512, 8><<<grid, 512, 0, s_>>>(bufs, fl, ep, rank, world, ov, nv); break;
    default: dp_allreduce_p2p_kernel<1024, 2><<<grid, 1024, 0, s_>>>(bufs, fl, ep, rank, world, ov, nv); break;
  }
  XV_CUDA_CHECK(cudaGetLastError());
  return XV_OK;
}
