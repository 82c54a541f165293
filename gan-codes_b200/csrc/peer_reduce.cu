// Data-parallel gradient exchange over NVLink peer memory (the reference has no multi-GPU path, its
// optimiser steps are deep_fusion_gan/model.py:195-197, 208-210, 226-228): ONE kernel all-reduces a range of the flat
// fp32 gradient buffer across the ranks of one node -- no NCCL call, no host synchronisation, capturable in the
// iteration's CUDA graph and small enough (no shared memory, <= 64 registers) to run next to the persistent conv
// kernels on a side stream.
//
// Every rank's gradient buffer and flag block live in cudaMalloc memory exported with cudaIpcGetMemHandle and mapped
// by the other ranks (dfgan_peer_* below), so a rank holds the address of every peer's buffer.
//
//   two-shot, in place:  rank r owns slice r of the range.  It (1) pulls slice r of every rank with 16-byte P2P loads,
//   adds them in rank order 0..world-1 (one reducer per slice, fixed order: every rank ends with bit-identical sums),
//   and (2) pushes the result into slice r of every rank with 16-byte P2P stores.  Slice r of a peer is read and written
//   by rank r only, so nothing orders (1) against (2).
//
//   flags: CTA b of rank r pairs with CTA b of every peer.  START: "my gradient is final" (being launched after the
//   producing kernels in stream order is what makes it true) is written into every peer's flag block with
//   st.release.sys and awaited with ld.acquire.sys; END: "slice r has landed in your buffer" likewise, after a
//   system-scope fence behind the pushes.  A rank that has seen END from every peer's CTA b for all b holds the whole
//   reduced range, and every peer has finished reading this rank's buffer -- the caller may overwrite it again.
//   The flag value is a per-CTA call counter kept in the local flag block (replays of a captured graph advance it).
//
// Sharded optimiser step (DFGAN_DP_SHARDED): the same kernel with the push phase removed is a reduce-scatter (rank r
// ends with the reduced slice r only); the caller then takes the squared norm of its slice, peer_sum_scalar_kernel sums
// the per-rank partials in rank order on every rank (one CTA, same flags), clip + Adam run on the owned slice only, and
// peer_allgather_kernel pushes the updated PARAMETER slice to every rank (END rendezvous): same bytes on the links,
// (N-1)/N of the clip + Adam passes gone from every rank.
//
// In-switch variant (peer_allreduce_mc_kernel): when the buffers are bound to an NVLS multicast object (the caller maps
// them, e.g. through torch's symmetric-memory allocator, and passes the multicast address of the range), rank r reduces
// slice r with ONE multimem.ld_reduce per 16 bytes -- the NVSwitch reads the element from every GPU and adds on the way
// -- and broadcasts the sum with ONE multimem.st: a GPU receives 1/N + (N-1)/N of the range instead of 2 (N-1)/N, and no
// SM adds anything.  Same flags, same slice ownership, every rank receives the single reducer's value (bit-identical).
//
// Waits are bounded by a poll count (a peer that died or deadlocked traps this kernel instead of hanging the box).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "host_common.h"

namespace dfgan {
namespace {

constexpr int kMaxWorld = 8;
constexpr int kMaxCtas = 128;
constexpr int kChannels = 6;
constexpr int kMaxThreads = 256;
constexpr int kLiteThreads = 128;
// flag block of one channel (uint32): [kMaxCtas][2 phases][kMaxWorld] flags, then [kMaxCtas] call counters
constexpr int kFlagWords = kMaxCtas * 2 * kMaxWorld + kMaxCtas;

struct PeerArgs {
  float4* buf[kMaxWorld];       // start of the range in rank p's buffer (own buffer at [rank])
  uint32_t* flag[kMaxWorld];    // rank p's flag block of this channel
  int world, rank;
  long long n4;                 // float4 elements in the range
  unsigned long long max_polls;   // bound of a wait on a peer, in polls of >= ~0.5 us each
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_peer(float4* p, const float4& v) {
  asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// threads 0..world-1 (except `rank`): signal peer t, then wait for peer t's signal in the local block
__device__ __forceinline__ void pair_barrier(const PeerArgs& a, int phase, uint32_t epoch) {
  const int t = threadIdx.x;
  if (t < a.world && t != a.rank) {
    const int slot = (blockIdx.x * 2 + phase) * kMaxWorld;
    st_release_sys(a.flag[t] + slot + a.rank, epoch);
    const uint32_t* mine = a.flag[a.rank] + slot + t;
    // The wait is bounded by COUNTING polls (one poll = a 64 ns sleep + a system-scope load, >= ~0.5 us), not by
    // %globaltimer: on these boxes the timer is re-synchronised while kernels run, and a difference of two reads then
    // under- or overflows -- every spinning CTA of a rank "timed out" at the same instant after a 3 ms wait
    // (round 2, 2 / 4 / 8 GPUs: `unspecified launch failure` in one of ~6 runs).
    unsigned long long polls = 0;
    // signed distance: the counter may wrap after 2^32 calls
    while (static_cast<int32_t>(ld_acquire_sys(mine) - epoch) < 0) {
      __nanosleep(64);
      if (++polls > a.max_polls) {
        printf("dfgan peer exchange: rank %d CTA %d gave up waiting for rank %d after %llu polls (phase %d, call %u)\n",
               a.rank, static_cast<int>(blockIdx.x), t, polls, phase, epoch);
        __trap();
      }
    }
  }
  __syncthreads();
}

// PC = peers loaded per round (loads in flight per thread = UNROLL * PC float4).  Two shapes are instantiated per world
// size: "full" (256 threads, up to 16 loads in flight: an exposed exchange, nothing else on the SMs) and "lite"
// (128 threads, <= 40 registers, 4 loads in flight: fits next to a resident conv_halo CTA, whose 320-352 threads x 168
// registers leave 6-11 K registers of an SM free -- the exchange that overlaps the backward sweep).
template <int WORLD, int UNROLL, int PC, int THREADS, int MIN_BLOCKS, bool SCATTER>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) peer_allreduce_kernel(const PeerArgs a) {
  uint32_t* ctr = a.flag[a.rank] + kMaxCtas * 2 * kMaxWorld + blockIdx.x;
  const uint32_t epoch = *ctr + 1u;
  pair_barrier(a, 0, epoch);

  const long long per = (a.n4 + WORLD - 1) / WORLD;
  const long long lo = per * a.rank;
  long long hi = lo + per;
  if (hi > a.n4) hi = a.n4;
  const int nthr = blockDim.x;
  const long long stride = static_cast<long long>(gridDim.x) * nthr * UNROLL;
  for (long long base = lo + static_cast<long long>(blockIdx.x) * nthr * UNROLL + threadIdx.x; base < hi;
       base += stride) {
    float4 s[UNROLL];
#pragma unroll
    for (int p0 = 0; p0 < WORLD; p0 += PC) {
      float4 v[UNROLL][PC];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const long long i = base + u * nthr;
        if (i < hi) {
#pragma unroll
          for (int p = 0; p < PC; ++p)
            if (p0 + p < WORLD) v[u][p] = ld_peer(a.buf[p0 + p] + i);
        }
      }
      // rank order 0..WORLD-1, whatever PC is
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
#pragma unroll
        for (int p = 0; p < PC; ++p) {
          if (p0 + p >= WORLD) continue;
          if (p0 + p == 0) {
            s[u] = v[u][0];
          } else {
            s[u].x += v[u][p].x;
            s[u].y += v[u][p].y;
            s[u].z += v[u][p].z;
            s[u].w += v[u][p].w;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long i = base + u * nthr;
      if (i < hi) {
        if (SCATTER) {
          st_peer(a.buf[a.rank] + i, s[u]);      // reduce-scatter: the owner keeps the sum, nothing is pushed
        } else {
#pragma unroll
          for (int p = 0; p < WORLD; ++p) st_peer(a.buf[p] + i, s[u]);
        }
      }
    }
  }
  if (!SCATTER) {
    __threadfence_system();
    __syncthreads();
    pair_barrier(a, 1, epoch);
  } else {
    __syncthreads();
  }
  if (threadIdx.x == 0) *ctr = epoch;
}

// rank r pushes slice r of the range (its freshly updated parameters) into every peer's buffer, then the END rendezvous
template <int UNROLL>
__global__ void __launch_bounds__(kMaxThreads, 2) peer_allgather_kernel(const PeerArgs a) {
  uint32_t* ctr = a.flag[a.rank] + kMaxCtas * 2 * kMaxWorld + blockIdx.x;
  const uint32_t epoch = *ctr + 1u;
  const long long per = (a.n4 + a.world - 1) / a.world;
  const long long lo = per * a.rank;
  long long hi = lo + per;
  if (hi > a.n4) hi = a.n4;
  const int nthr = blockDim.x;
  const long long stride = static_cast<long long>(gridDim.x) * nthr * UNROLL;
  const float4* mine = a.buf[a.rank];
  for (long long base = lo + static_cast<long long>(blockIdx.x) * nthr * UNROLL + threadIdx.x; base < hi;
       base += stride) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long i = base + u * nthr;
      if (i < hi) v[u] = mine[i];
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long i = base + u * nthr;
      if (i < hi) {
        for (int p = 0; p < a.world; ++p)
          if (p != a.rank) st_peer(a.buf[p] + i, v[u]);
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  pair_barrier(a, 1, epoch);
  if (threadIdx.x == 0) *ctr = epoch;
}

// total = sum over ranks (in rank order, identical on every rank) of one float per rank; a.buf[p] = rank p's slot array
// (kMaxWorld floats); one CTA of 32 threads
__global__ void __launch_bounds__(32) peer_sum_scalar_kernel(const PeerArgs a, const float* partial, float* total) {
  uint32_t* ctr = a.flag[a.rank] + kMaxCtas * 2 * kMaxWorld;
  const uint32_t epoch = *ctr + 1u;
  const int t = threadIdx.x;
  if (t < a.world) {
    float* dst = reinterpret_cast<float*>(a.buf[t]) + a.rank;
    const float v = *partial;
    asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst), "f"(v) : "memory");
  }
  pair_barrier(a, 0, epoch);       // the release store of thread t follows its data store to rank t
  if (t == 0) {
    float s = 0.f;
    for (int p = 0; p < a.world; ++p) {
      float v;
      asm volatile("ld.relaxed.sys.global.f32 %0, [%1];"
                   : "=f"(v)
                   : "l"(reinterpret_cast<float*>(a.buf[a.rank]) + p)
                   : "memory");
      s += v;
    }
    *total = s;
    *ctr = epoch;
  }
}

__device__ __forceinline__ float4 multimem_ld_reduce(const float4* mc) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(mc)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float4* mc, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

template <int UNROLL, int THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS) peer_allreduce_mc_kernel(const PeerArgs a, float4* mc) {
  uint32_t* ctr = a.flag[a.rank] + kMaxCtas * 2 * kMaxWorld + blockIdx.x;
  const uint32_t epoch = *ctr + 1u;
  pair_barrier(a, 0, epoch);

  const long long per = (a.n4 + a.world - 1) / a.world;
  const long long lo = per * a.rank;
  long long hi = lo + per;
  if (hi > a.n4) hi = a.n4;
  const int nthr = blockDim.x;
  const long long stride = static_cast<long long>(gridDim.x) * nthr * UNROLL;
  for (long long base = lo + static_cast<long long>(blockIdx.x) * nthr * UNROLL + threadIdx.x; base < hi;
       base += stride) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long i = base + u * nthr;
      if (i < hi) v[u] = multimem_ld_reduce(mc + i);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const long long i = base + u * nthr;
      if (i < hi) multimem_st(mc + i, v[u]);
    }
  }
  __threadfence_system();
  __syncthreads();
  pair_barrier(a, 1, epoch);
  if (threadIdx.x == 0) *ctr = epoch;
}

template <int WORLD, bool SCATTER>
cudaError_t launch_world(const PeerArgs& a, int ctas, int threads, cudaStream_t st) {
  if (threads > kLiteThreads) {
    constexpr int U = WORLD <= 2 ? 8 : (WORLD <= 4 ? 4 : 2);
    peer_allreduce_kernel<WORLD, U, WORLD, kMaxThreads, 2, SCATTER><<<ctas, threads, 0, st>>>(a);
  } else {
    constexpr int PC = WORLD < 4 ? WORLD : 4;
    constexpr int U = WORLD <= 2 ? 2 : 1;
    peer_allreduce_kernel<WORLD, U, PC, kLiteThreads, 12, SCATTER><<<ctas, threads, 0, st>>>(a);
  }
  return cudaGetLastError();
}

}  // namespace
}  // namespace dfgan

using namespace dfgan;

static int peer_prepare(PeerArgs& a, void* const* flags, int world, int rank, int64_t count, int channel, int& ctas,
                        int& threads, int timeout_ms) {
  if (!flags || world < 2 || world > kMaxWorld || rank < 0 || rank >= world || count <= 0 || (count & 3) ||
      channel < 0 || channel >= kChannels)
    return DFGAN_E_ARG;
  if (ctas <= 0) ctas = 32;
  if (ctas > kMaxCtas) ctas = kMaxCtas;
  for (int p = 0; p < world; ++p) {
    if (!flags[p]) return DFGAN_E_ARG;
    a.flag[p] = static_cast<uint32_t*>(flags[p]) + static_cast<size_t>(channel) * kFlagWords;
  }
  a.world = world, a.rank = rank, a.n4 = count / 4;
  a.max_polls = static_cast<unsigned long long>(timeout_ms > 0 ? timeout_ms : 20000) * 1000ull;   // ~1 us per poll
  if (threads <= 0) threads = kMaxThreads;
  threads = (threads / 32) * 32;
  threads = threads < 32 ? 32 : (threads > kMaxThreads ? kMaxThreads : threads);
  // few elements: fewer CTAs (every CTA pays two flag round trips); the same on every rank
  const long long per = (a.n4 + world - 1) / world;
  const long long want = (per + threads * 4 - 1) / (threads * 4);
  if (want < ctas) ctas = static_cast<int>(want < 1 ? 1 : want);
  return 0;
}

template <bool SCATTER>
static int peer_reduce_launch(void* const* bufs, void* const* flags, int world, int rank, int64_t count, int channel,
                              int ctas, int threads, int timeout_ms, dfgan_stream_t stream) {
  PeerArgs a{};
  if (!bufs) return DFGAN_E_ARG;
  const int rc = peer_prepare(a, flags, world, rank, count, channel, ctas, threads, timeout_ms);
  if (rc) return rc;
  for (int p = 0; p < world; ++p) {
    if (!bufs[p] || (reinterpret_cast<uintptr_t>(bufs[p]) & 15)) return DFGAN_E_ARG;
    a.buf[p] = static_cast<float4*>(bufs[p]);
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e;
  switch (world) {
    case 2: e = launch_world<2, SCATTER>(a, ctas, threads, st); break;
    case 3: e = launch_world<3, SCATTER>(a, ctas, threads, st); break;
    case 4: e = launch_world<4, SCATTER>(a, ctas, threads, st); break;
    case 5: e = launch_world<5, SCATTER>(a, ctas, threads, st); break;
    case 6: e = launch_world<6, SCATTER>(a, ctas, threads, st); break;
    case 7: e = launch_world<7, SCATTER>(a, ctas, threads, st); break;
    default: e = launch_world<8, SCATTER>(a, ctas, threads, st); break;
  }
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  return 0;
}

extern "C" {

size_t dfgan_peer_flag_bytes(void) { return static_cast<size_t>(kChannels) * kFlagWords * sizeof(uint32_t); }
int dfgan_peer_handle_bytes(void) { return static_cast<int>(sizeof(cudaIpcMemHandle_t)); }

int dfgan_peer_alloc(size_t bytes, void** ptr) {
  if (!ptr || bytes == 0) return DFGAN_E_ARG;
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaMemset(*ptr, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  return static_cast<int>(e);
}

int dfgan_peer_free(void* ptr) { return static_cast<int>(cudaFree(ptr)); }

int dfgan_peer_export(void* ptr, void* handle_out) {
  if (!ptr || !handle_out) return DFGAN_E_ARG;
  return static_cast<int>(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle_out), ptr));
}

int dfgan_peer_import(const void* handle, void** ptr) {
  if (!handle || !ptr) return DFGAN_E_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  return static_cast<int>(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
}

int dfgan_peer_close(void* ptr) { return static_cast<int>(cudaIpcCloseMemHandle(ptr)); }

int dfgan_peer_allreduce_f32(void* const* bufs, void* const* flags, int world, int rank, int64_t count, int channel,
                             int ctas, int threads, int timeout_ms, dfgan_stream_t stream) {
  return peer_reduce_launch<false>(bufs, flags, world, rank, count, channel, ctas, threads, timeout_ms, stream);
}

int dfgan_peer_reduce_scatter_f32(void* const* bufs, void* const* flags, int world, int rank, int64_t count,
                                  int channel, int ctas, int threads, int timeout_ms, dfgan_stream_t stream) {
  return peer_reduce_launch<true>(bufs, flags, world, rank, count, channel, ctas, threads, timeout_ms, stream);
}

int dfgan_peer_allgather_f32(void* const* bufs, void* const* flags, int world, int rank, int64_t count, int channel,
                             int ctas, int timeout_ms, dfgan_stream_t stream) {
  PeerArgs a{};
  if (!bufs) return DFGAN_E_ARG;
  int threads = kMaxThreads;
  const int rc = peer_prepare(a, flags, world, rank, count, channel, ctas, threads, timeout_ms);
  if (rc) return rc;
  for (int p = 0; p < world; ++p) {
    if (!bufs[p] || (reinterpret_cast<uintptr_t>(bufs[p]) & 15)) return DFGAN_E_ARG;
    a.buf[p] = static_cast<float4*>(bufs[p]);
  }
  peer_allgather_kernel<4><<<ctas, threads, 0, static_cast<cudaStream_t>(stream)>>>(a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  return 0;
}

int dfgan_peer_sum_scalar_f32(const float* partial, void* const* slots, float* total, void* const* flags, int world,
                              int rank, int channel, int timeout_ms, dfgan_stream_t stream) {
  PeerArgs a{};
  if (!partial || !slots || !total) return DFGAN_E_ARG;
  int ctas = 1, threads = 32;
  const int rc = peer_prepare(a, flags, world, rank, 4, channel, ctas, threads, timeout_ms);
  if (rc) return rc;
  for (int p = 0; p < world; ++p) {
    if (!slots[p] || (reinterpret_cast<uintptr_t>(slots[p]) & 15)) return DFGAN_E_ARG;
    a.buf[p] = static_cast<float4*>(slots[p]);
  }
  peer_sum_scalar_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(a, partial, total);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  return 0;
}

int dfgan_peer_allreduce_mc_f32(void* mc_range, void* const* flags, int world, int rank, int64_t count, int channel,
                                int ctas, int threads, int timeout_ms, dfgan_stream_t stream) {
  PeerArgs a{};
  if (!mc_range || (reinterpret_cast<uintptr_t>(mc_range) & 15)) return DFGAN_E_ARG;
  const int rc = peer_prepare(a, flags, world, rank, count, channel, ctas, threads, timeout_ms);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float4* mc = static_cast<float4*>(mc_range);
  if (threads > kLiteThreads)
    peer_allreduce_mc_kernel<8, kMaxThreads, 2><<<ctas, threads, 0, st>>>(a, mc);
  else
    peer_allreduce_mc_kernel<4, kLiteThreads, 12><<<ctas, threads, 0, st>>>(a, mc);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return static_cast<int>(e);
  count_launch();
  return 0;
}

}  // extern "C"
