// Gradient exchange of the data-parallel train step over NVLink 5 / NVSwitch peer memory (one SUM all-reduce of
// the flat gradient buffer per step, train.py:111-113 on every replica afterwards).
//
// The flat fp32 gradient range is packed to bf16 into a SYMMETRIC staging buffer (same allocation on every rank, mapped into
// every peer and -- where the fabric supports it -- bound to one NVSwitch multicast address).  Then
//   dp_reduce_bcast_kernel : rank r owns slice r of the range.  With multicast it issues multimem.ld_reduce (the switch adds
//                            the R staging copies in fp32 and returns bf16x2) and multimem.st (the switch writes the sum back
//                            into all R copies): every byte crosses this GPU's links once in and once out, and the reduction
//                            itself costs no SM arithmetic.  Without multicast it reads the slice from every peer pointer,
//                            adds in fp32 and stores the result to every peer pointer.
//   dp_wait_unpack_kernel  : waits until every rank has broadcast its slice, then unpacks the whole range back to fp32.
// Cross-GPU ordering uses two flags per (rank, peer) in a second symmetric buffer, written with st.release.sys after a
// system-scope fence and polled with ld.acquire.sys; the value is a per-call epoch, so nothing is ever reset.  Kernel
// boundaries on the communication stream provide the intra-GPU ordering (pack -> reduce -> unpack), so no grid-wide barrier
// is needed and a block that is scheduled late only delays itself.  Every wait is bounded (20 s) and traps.
#include <algorithm>

#include "common.cuh"

namespace regat {
namespace {

constexpr int MAX_RANKS = 16;
constexpr unsigned long long WAIT_NS = 20000000000ull;   // 20 s: ranks may be seconds apart while graphs are being captured

__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// flags layout in every rank's buffer: [kind][source rank], kind 0 = "my range is packed", 1 = "my slice is broadcast"
__device__ __forceinline__ void signal_all(unsigned int* const* flag_ptrs, int kind, int rank, int world, unsigned int epoch) {
  __threadfence_system();
  if ((int)threadIdx.x < world) st_release_sys(flag_ptrs[threadIdx.x] + kind * MAX_RANKS + rank, epoch);
}
__device__ __forceinline__ void wait_all(const unsigned int* my_flags, int kind, int world, unsigned int epoch) {
  if ((int)threadIdx.x < world) {
    const unsigned int* f = my_flags + kind * MAX_RANKS + threadIdx.x;
    const unsigned long long t0 = globaltimer_ns();
    // epochs only grow; the signed difference tolerates wrap-around
    while ((int)(ld_acquire_sys(f) - epoch) < 0) {
      if (globaltimer_ns() - t0 > WAIT_NS) __trap();   // a lost peer must surface as an error, never as a hung GPU
    }
  }
  __syncthreads();
}

struct PeerPtrs { void* p[MAX_RANKS]; };

template <bool MULTICAST>
__global__ void __launch_bounds__(512) dp_reduce_bcast_kernel(PeerPtrs stage, unsigned char* mc, PeerPtrs flags, int rank, int world,
                                                               long long offset, long long numel, unsigned int epoch) {
  __shared__ unsigned int* fl[MAX_RANKS];
  if ((int)threadIdx.x < world) fl[threadIdx.x] = static_cast<unsigned int*>(flags.p[threadIdx.x]);
  __syncthreads();
  if (blockIdx.x == 0) signal_all(fl, 0, rank, world, epoch);          // the pack kernel before this one has completed
  wait_all(fl[rank], 0, world, epoch);                                  // ... on every rank
  // slice of this rank, in 16-byte (8 x bf16) chunks
  const long long chunks = (numel + 7) / 8, per = (chunks + world - 1) / world;
  const long long c0 = (long long)rank * per, c1 = min(chunks, c0 + per);
  const long long stride = (long long)gridDim.x * blockDim.x;
  if constexpr (MULTICAST) {
    // four reductions in flight per thread: the switch round trip (a few microseconds) is the latency to hide
    constexpr int U = 4;
    for (long long c = c0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; c < c1; c += U * stride) {
      unsigned int v[U][4];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long cu = c + u * stride;
        if (cu < c1)
          asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
                       : "=r"(v[u][0]), "=r"(v[u][1]), "=r"(v[u][2]), "=r"(v[u][3])
                       : "l"(mc + (offset + cu * 8) * 2)
                       : "memory");
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long cu = c + u * stride;
        if (cu < c1)
          asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(mc + (offset + cu * 8) * 2), "r"(v[u][0]),
                       "r"(v[u][1]), "r"(v[u][2]), "r"(v[u][3])
                       : "memory");
      }
    }
  } else {
    for (long long c = c0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; c < c1; c += stride) {
      const long long byte = (offset + c * 8) * 2;
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int r = 0; r < world; ++r) {
        const uint4 x = __ldcv(reinterpret_cast<const uint4*>(static_cast<unsigned char*>(stage.p[r]) + byte));
        const unsigned int w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[2 * i] += __uint_as_float(w[i] << 16); acc[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u); }
      }
      unsigned int o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]);
        o[i] = *reinterpret_cast<unsigned int*>(&h);
      }
      for (int r = 0; r < world; ++r)
        *reinterpret_cast<uint4*>(static_cast<unsigned char*>(stage.p[r]) + byte) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  __threadfence_system();
}

__global__ void __launch_bounds__(256) dp_wait_unpack_kernel(const bf16* __restrict__ stage_local, float* __restrict__ dst, PeerPtrs flags,
                                                             int rank, int world, long long offset, long long numel, unsigned int epoch) {
  __shared__ unsigned int* fl[MAX_RANKS];
  if ((int)threadIdx.x < world) fl[threadIdx.x] = static_cast<unsigned int*>(flags.p[threadIdx.x]);
  __syncthreads();
  if (blockIdx.x == 0) signal_all(fl, 1, rank, world, epoch);          // my reduce/broadcast kernel has completed
  wait_all(fl[rank], 1, world, epoch);                                  // every slice of the range has landed in my staging copy
  const long long n8 = numel / 8;
  const uint4* src = reinterpret_cast<const uint4*>(stage_local + offset);
  float4* out = reinterpret_cast<float4*>(dst);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
    // ld.cv: written by peers / the switch, never served from a stale cache line
    const uint4 x = __ldcv(src + i);
    out[2 * i] = make_float4(__uint_as_float(x.x << 16), __uint_as_float(x.x & 0xffff0000u), __uint_as_float(x.y << 16),
                             __uint_as_float(x.y & 0xffff0000u));
    out[2 * i + 1] = make_float4(__uint_as_float(x.z << 16), __uint_as_float(x.z & 0xffff0000u), __uint_as_float(x.w << 16),
                                 __uint_as_float(x.w & 0xffff0000u));
  }
}

// fp32 variant, in place on a SYMMETRIC gradient buffer: no staging copy and no pack / unpack passes over HBM -- the only
// traffic is the reduction itself.  Twice the link bytes of the bf16 wire format, which NVLink 5 absorbs (76 MB per step and
// GPU in each direction at 8 ranks); what it buys is zero extra kernels beside the backward pass and unrounded gradients.
template <bool MULTICAST>
__global__ void __launch_bounds__(512) dp_reduce_bcast_f32_kernel(PeerPtrs grads, unsigned char* mc, PeerPtrs flags, int rank, int world,
                                                                   long long offset, long long numel, unsigned int epoch,
                                                                   const unsigned int* epoch_dev) {
  __shared__ unsigned int* fl[MAX_RANKS];
  if (epoch_dev) epoch = *epoch_dev + 1u;      // device-resident call counter (CUDA-graph replay): advanced by dp_wait_kernel
  if ((int)threadIdx.x < world) fl[threadIdx.x] = static_cast<unsigned int*>(flags.p[threadIdx.x]);
  __syncthreads();
  if (blockIdx.x == 0) signal_all(fl, 0, rank, world, epoch);          // the kernels that produced this range have completed
  wait_all(fl[rank], 0, world, epoch);                                  // ... on every rank
  const long long chunks = numel / 4, per = (chunks + world - 1) / world;       // 16-byte (4 x fp32) chunks
  const long long c0 = (long long)rank * per, c1 = min(chunks, c0 + per);
  const long long stride = (long long)gridDim.x * blockDim.x;
  if constexpr (MULTICAST) {
    constexpr int U = 4;
    for (long long c = c0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; c < c1; c += U * stride) {
      float v[U][4];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long cu = c + u * stride;
        if (cu < c1)
          asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(v[u][0]), "=f"(v[u][1]), "=f"(v[u][2]), "=f"(v[u][3])
                       : "l"(mc + (offset + cu * 4) * 4)
                       : "memory");
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long cu = c + u * stride;
        if (cu < c1)
          asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + (offset + cu * 4) * 4), "f"(v[u][0]),
                       "f"(v[u][1]), "f"(v[u][2]), "f"(v[u][3])
                       : "memory");
      }
    }
  } else {
    for (long long c = c0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; c < c1; c += stride) {
      const long long byte = (offset + c * 4) * 4;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < world; ++r) {
        const float4 x = __ldcv(reinterpret_cast<const float4*>(static_cast<unsigned char*>(grads.p[r]) + byte));
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
      }
      for (int r = 0; r < world; ++r) *reinterpret_cast<float4*>(static_cast<unsigned char*>(grads.p[r]) + byte) = acc;
    }
  }
  __threadfence_system();
}

// one block: "my slice is broadcast" to every peer, then wait for every peer's -- after it the whole range is final here
__global__ void __launch_bounds__(32) dp_wait_kernel(PeerPtrs flags, int rank, int world, unsigned int epoch, unsigned int* epoch_dev) {
  __shared__ unsigned int* fl[MAX_RANKS];
  if (epoch_dev) epoch = *epoch_dev + 1u;
  if ((int)threadIdx.x < world) fl[threadIdx.x] = static_cast<unsigned int*>(flags.p[threadIdx.x]);
  __syncthreads();
  signal_all(fl, 1, rank, world, epoch);
  wait_all(fl[rank], 1, world, epoch);
  if (epoch_dev && threadIdx.x == 0) *epoch_dev = epoch;      // the next exchange on this stream uses epoch + 1
}

int fill(PeerPtrs& pp, const uint64_t* host_ptrs, int world) {
  for (int r = 0; r < MAX_RANKS; ++r) pp.p[r] = r < world ? reinterpret_cast<void*>(host_ptrs[r]) : nullptr;
  return REGAT_OK;
}

}  // namespace
}  // namespace regat

using namespace regat;

extern "C" int regat_dp_reduce_bcast(const uint64_t* stage_ptrs, uint64_t multicast_ptr, const uint64_t* flag_ptrs, int rank, int world,
                                     int64_t offset, int64_t numel, uint32_t epoch, int blocks, regat_stream_t stream) {
  REGAT_REQUIRE(stage_ptrs && flag_ptrs, REGAT_ERR_ARG, "dp_reduce_bcast: null pointer table");
  REGAT_REQUIRE(world >= 1 && world <= MAX_RANKS && rank >= 0 && rank < world, REGAT_ERR_ARG, "dp_reduce_bcast: bad rank %d / world %d", rank, world);
  REGAT_REQUIRE(offset % 8 == 0 && numel % 8 == 0 && numel >= 0, REGAT_ERR_ALIGN, "dp_reduce_bcast: offset and count must be multiples of 8 elements");
  if (numel == 0) return REGAT_OK;
  PeerPtrs st, fl;
  fill(st, stage_ptrs, world); fill(fl, flag_ptrs, world);
  const int nb = blocks > 0 ? blocks : 32;
  cudaStream_t s = (cudaStream_t)stream;
  if (multicast_ptr)
    dp_reduce_bcast_kernel<true><<<nb, 512, 0, s>>>(st, reinterpret_cast<unsigned char*>(multicast_ptr), fl, rank, world, offset, numel, epoch);
  else
    dp_reduce_bcast_kernel<false><<<nb, 512, 0, s>>>(st, nullptr, fl, rank, world, offset, numel, epoch);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}

extern "C" int regat_dp_wait_unpack(const void* stage_local, float* dst, const uint64_t* flag_ptrs, int rank, int world, int64_t offset,
                                    int64_t numel, uint32_t epoch, regat_stream_t stream) {
  REGAT_REQUIRE(stage_local && dst && flag_ptrs, REGAT_ERR_ARG, "dp_wait_unpack: null pointer");
  REGAT_REQUIRE(world >= 1 && world <= MAX_RANKS && rank >= 0 && rank < world, REGAT_ERR_ARG, "dp_wait_unpack: bad rank %d / world %d", rank, world);
  REGAT_REQUIRE(offset % 8 == 0 && numel % 8 == 0 && aligned16(dst), REGAT_ERR_ALIGN, "dp_wait_unpack: range must be 8-element / 16-byte aligned");
  if (numel == 0) return REGAT_OK;
  PeerPtrs fl;
  fill(fl, flag_ptrs, world);
  const int nb = (int)std::min<long long>((numel / 8 + 255) / 256, (long long)num_sms() * 2);
  dp_wait_unpack_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(static_cast<const bf16*>(stage_local), dst, fl, rank, world, offset, numel, epoch);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}

namespace regat {
int dp_allreduce_f32_impl(const uint64_t* grad_ptrs, uint64_t multicast_ptr, const uint64_t* flag_ptrs, int rank, int world,
                          int64_t offset, int64_t numel, uint32_t epoch, uint32_t* epoch_dev, int blocks, cudaStream_t stream);
}
extern "C" int regat_dp_allreduce_f32(const uint64_t* grad_ptrs, uint64_t multicast_ptr, const uint64_t* flag_ptrs, int rank, int world,
                                      int64_t offset, int64_t numel, uint32_t epoch, int blocks, regat_stream_t stream) {
  return dp_allreduce_f32_impl(grad_ptrs, multicast_ptr, flag_ptrs, rank, world, offset, numel, epoch, nullptr, blocks, (cudaStream_t)stream);
}
extern "C" int regat_dp_allreduce_f32_dev(const uint64_t* grad_ptrs, uint64_t multicast_ptr, const uint64_t* flag_ptrs, int rank, int world,
                                          int64_t offset, int64_t numel, uint32_t* epoch_dev, int blocks, regat_stream_t stream) {
  REGAT_REQUIRE(epoch_dev, REGAT_ERR_ARG, "dp_allreduce_f32_dev: null epoch counter");
  return dp_allreduce_f32_impl(grad_ptrs, multicast_ptr, flag_ptrs, rank, world, offset, numel, 0, epoch_dev, blocks, (cudaStream_t)stream);
}
int regat::dp_allreduce_f32_impl(const uint64_t* grad_ptrs, uint64_t multicast_ptr, const uint64_t* flag_ptrs, int rank, int world,
                                 int64_t offset, int64_t numel, uint32_t epoch, uint32_t* epoch_dev, int blocks, cudaStream_t stream) {
  REGAT_REQUIRE(grad_ptrs && flag_ptrs, REGAT_ERR_ARG, "dp_allreduce_f32: null pointer table");
  REGAT_REQUIRE(world >= 1 && world <= MAX_RANKS && rank >= 0 && rank < world, REGAT_ERR_ARG, "dp_allreduce_f32: bad rank %d / world %d", rank, world);
  REGAT_REQUIRE(offset % 4 == 0 && numel % 4 == 0 && numel >= 0, REGAT_ERR_ALIGN, "dp_allreduce_f32: offset and count must be multiples of 4 elements");
  if (numel == 0) return REGAT_OK;
  PeerPtrs gp, fl;
  fill(gp, grad_ptrs, world); fill(fl, flag_ptrs, world);
  // 32 x 512 threads measured best at 2 and 8 ranks: enough reductions in flight to cover the switch round trip, few enough
  // CTAs to leave the backward pass its SMs (64 x 256 threads at 40 registers, which would fit beside a resident GEMM CTA, was
  // slower: the register cap costs the memory-level parallelism the kernel lives on).
  const int nb = blocks > 0 ? blocks : 32;
  cudaStream_t s = (cudaStream_t)stream;
  if (multicast_ptr)
    dp_reduce_bcast_f32_kernel<true><<<nb, 512, 0, s>>>(gp, reinterpret_cast<unsigned char*>(multicast_ptr), fl, rank, world, offset, numel, epoch, epoch_dev);
  else
    dp_reduce_bcast_f32_kernel<false><<<nb, 512, 0, s>>>(gp, nullptr, fl, rank, world, offset, numel, epoch, epoch_dev);
  REGAT_POST_LAUNCH();
  dp_wait_kernel<<<1, 32, 0, s>>>(fl, rank, world, epoch, epoch_dev);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
