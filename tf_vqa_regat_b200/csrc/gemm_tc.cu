// Kernel (b): bf16 dense projections on the 5th-gen tensor cores.
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring -> tcgen05.mma (one elected thread,
//   fp32 accumulators in TMEM) -> tcgen05.ld epilogue with the fused regat_epilogue.
// Replaces the reference's tf.keras Dense calls and their autodiff transposes (fc.py:36-43,
// graph_att_layer.py:47,55,117, fusion.py:32-39, classifier.py:14-19; SURVEY K2,K4,K5,K11,K13,K14).
//
// One kernel covers forward (A K-major, B MN-major), dgrad (both K-major) and wgrad (both MN-major)
// without any transposed copy in HBM: the operand's major-ness only changes the TMA box shape and the
// shared-memory matrix descriptor (UMMA canonical layouts, SWIZZLE_128B):
//   K-major  tile: rows x 64 bf16 (128 B per row), 8-row swizzle atoms: SBO = 1024 B, +32 B per UMMA_K
//   MN-major tile: 64-element MN chunks of [64 K-rows x 128 B]: LBO = 8192 B (next chunk), SBO = 1024 B
//                  (next 8 K-rows), +2048 B per UMMA_K
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2-5 = epilogue
// (warp w reads TMEM lanes 32*(w%4)..).  Every mbarrier wait is bounded and traps instead of hanging.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace regat {
namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int NTHREADS = 192;
constexpr uint32_t SPIN_LIMIT = 1u << 27;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < SPIN_LIMIT; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
  }
  __trap();   // a pipeline bug must surface as a CUDA error, never as a hung GPU
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor (tcgen05 / "UMMA"): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout SWIZZLE_128B = 2 at [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, a_major bit15, b_major bit16,
// N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(BM >> 4) << 24);
}

struct TcParams {
  int M, N, K, ldc, c_f32, k_blocks_per_split, total_k_blocks, atomic_out, tiles_m, tiles_n, splits;
  int c_block_cols;               // > 0: column block j = c / c_block_cols of C lives at C + c_block_off[j] (fp32 outputs only)
  long long c_block_off[8];
  void* C;
  EpiArgs e;
};

// Persistent: CTA c processes work units c, c+grid, ... where a unit = (m-tile, n-tile, k-split), n fastest so that
// concurrently running CTAs share A rows in L2.  ACC accumulator stages in TMEM (ACC*BN <= 512 columns) let the epilogue
// of unit i overlap the MMAs of unit i+1.
template <int BN, int STAGES, int ACC, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NTHREADS) gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA,
                                                           const __grid_constant__ CUtensorMap mapB, const TcParams p) {
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = ACC * BN;
  static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two");
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;       // [ACC]
  uint64_t* tmem_empty = tmem_full + ACC;         // [ACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + ACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = p.tiles_n, splits = p.splits;
  const int units = p.tiles_m * tiles_n * splits;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int a = 0; a < ACC; ++a) { mbar_init(tmem_full + a, 1); mbar_init(tmem_empty + a, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // unit -> (m0, n0, kb0, kb1)
  auto decode = [&](int u, int& m0, int& n0, int& kb0, int& kb1) {
    const int tile = u / splits, z = u - tile * splits;
    const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
    m0 = tm * BM; n0 = tn * BN;
    kb0 = z * p.k_blocks_per_split;
    kb1 = min(p.total_k_blocks, kb0 + p.k_blocks_per_split);
  };

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int m0, n0, kb0, kb1;
        decode(u, m0, n0, kb0, kb1);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(empty_bar + s, ph ^ 1);
          mbar_expect_tx(full_bar + s, STAGE_BYTES);
          unsigned char* sa = tiles + s * STAGE_BYTES;
          unsigned char* sb = sa + A_BYTES;
          if (!A_MN) {
            tma_load_2d(sa, &mapA, full_bar + s, kb * BK, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * (64 * BK * 2), &mapA, full_bar + s, m0 + c * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(sb, &mapB, full_bar + s, kb * BK, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * (64 * BK * 2), &mapB, full_bar + s, n0 + c * 64, kb * BK);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BN, A_MN, B_MN);
      int it = 0, ui = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x, ++ui) {
        int m0, n0, kb0, kb1;
        decode(u, m0, n0, kb0, kb1);
        const int a = ui % ACC;
        const uint32_t aph = (ui / ACC) & 1;
        mbar_wait(tmem_empty + a, aph ^ 1);          // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(a * BN);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(full_bar + s, ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + s * STAGE_BYTES), sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = A_MN ? make_desc(sa + k * 2048, 64 * BK * 2, 1024) : make_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_desc(sb + k * 2048, 64 * BK * 2, 1024) : make_desc(sb + k * 32, 16, 1024);
            umma_bf16(tmem_d, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar + s);          // frees the smem slot once these MMAs have read it
        }
        umma_commit(tmem_full + a);            // accumulator of this unit complete
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> fused epilogue -> global =====
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    int ui = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++ui) {
    int m0, n0, kb0, kb1;
    decode(u, m0, n0, kb0, kb1);
    const int acc_stage = ui % ACC;
    mbar_wait(tmem_full + acc_stage, (ui / ACC) & 1);
    tc_fence_after();
    const uint32_t tmem_acc = tmem_base + (uint32_t)(acc_stage * BN);
    const int r = m0 + q * 32 + lane;
    const bool row_ok = r < p.M && kb1 > kb0;
    const int rc = min(r, p.M - 1);
    // per-row epilogue terms, hoisted out of the column loop
    const float rs = (p.e.addend && p.e.row_scale) ? __ldg(p.e.row_scale + rc) : 1.f;
    const float* arow = p.e.addend ? p.e.addend + (size_t)(rc / p.e.addend_rows) * p.e.addend_ld : nullptr;
    const float alpha0 = (p.e.alpha && p.e.alpha_cols == 0) ? __ldg(p.e.alpha) : 1.f;
    int r2 = -1;
    if (p.e.c2) {
      const int rr = rc % p.e.c2_rows_in;
      if (rr < p.e.c2_rows_keep) r2 = (rc / p.e.c2_rows_in) * p.e.c2_rows_keep + rr;
    }
    // fast path needs 16-byte aligned rows for every tensor touched with vector accesses
    const size_t esz = p.c_f32 ? 4 : 2;
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p.C) | ((size_t)p.ldc * esz)) & 15) == 0 &&
                        (!p.e.bias || (reinterpret_cast<uintptr_t>(p.e.bias) & 15) == 0) &&
                        (!p.e.addend || ((reinterpret_cast<uintptr_t>(p.e.addend) | ((size_t)p.e.addend_ld * 4)) & 15) == 0) &&
                        (!p.e.gate || ((reinterpret_cast<uintptr_t>(p.e.gate) | ((size_t)p.e.gate_ld * esz)) & 15) == 0) &&
                        (!p.e.c2 || ((reinterpret_cast<uintptr_t>(p.e.c2) | ((size_t)p.e.c2_ld * esz)) & 15) == 0) &&
                        (p.e.alpha_cols % 32 == 0);
#pragma unroll 1
    for (int cc = 0; cc < BN / 32; ++cc) {
      float v[32];
      tmem_ld32(tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(cc * 32), v);
      const int c0 = n0 + cc * 32;
      if (!row_ok || c0 >= p.N) continue;
      if (p.atomic_out) {
        float* C = static_cast<float*>(p.C);
        int cl = c0;
        if (p.c_block_cols) { const int jb = c0 / p.c_block_cols; C += p.c_block_off[jb]; cl = c0 - jb * p.c_block_cols; }
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j < p.N) {
            float x = v[j];
            if (p.e.alpha) x *= p.e.alpha[p.e.alpha_cols ? (c0 + j) / p.e.alpha_cols : 0];
            atomicAdd(C + (size_t)r * p.ldc + cl + j, x);
          }
      } else if (vec_ok && c0 + 32 <= p.N) {
        // ---- vectorised, branch-free path: every load is issued before its first use
        if (arow) {
          const float4* ap = reinterpret_cast<const float4*>(arow + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 a = __ldg(ap + j);
            v[4 * j] = fmaf(rs, a.x, v[4 * j]); v[4 * j + 1] = fmaf(rs, a.y, v[4 * j + 1]);
            v[4 * j + 2] = fmaf(rs, a.z, v[4 * j + 2]); v[4 * j + 3] = fmaf(rs, a.w, v[4 * j + 3]);
          }
        }
        const float al = p.e.alpha ? (p.e.alpha_cols ? __ldg(p.e.alpha + c0 / p.e.alpha_cols) : alpha0) : 1.f;
        if (p.e.bias) {
          const float4* bp = reinterpret_cast<const float4*>(p.e.bias + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = __ldg(bp + j);
            v[4 * j] = fmaf(v[4 * j], al, bb.x); v[4 * j + 1] = fmaf(v[4 * j + 1], al, bb.y);
            v[4 * j + 2] = fmaf(v[4 * j + 2], al, bb.z); v[4 * j + 3] = fmaf(v[4 * j + 3], al, bb.w);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= al;
        }
        if (p.e.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (p.c_f32) {
          float* dst = static_cast<float*>(p.C) + (size_t)r * p.ldc + c0;
          if (p.c_block_cols) {
            const int jb = c0 / p.c_block_cols;
            dst = static_cast<float*>(p.C) + p.c_block_off[jb] + (size_t)r * p.ldc + (c0 - jb * p.c_block_cols);
          }
          if (p.e.accumulate) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 o = *reinterpret_cast<const float4*>(dst + 4 * j);
              v[4 * j] += o.x; v[4 * j + 1] += o.y; v[4 * j + 2] += o.z; v[4 * j + 3] += o.w;
            }
          }
          if (p.e.gate) {
            const float4* gp = reinterpret_cast<const float4*>(static_cast<const float*>(p.e.gate) + (size_t)r * p.e.gate_ld + c0);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 g4 = __ldg(gp + j);
              v[4 * j] = g4.x > 0.f ? v[4 * j] : 0.f; v[4 * j + 1] = g4.y > 0.f ? v[4 * j + 1] : 0.f;
              v[4 * j + 2] = g4.z > 0.f ? v[4 * j + 2] : 0.f; v[4 * j + 3] = g4.w > 0.f ? v[4 * j + 3] : 0.f;
            }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          if (r2 >= 0) {
            float* d2 = static_cast<float*>(p.e.c2) + (size_t)r2 * p.e.c2_ld + c0;
#pragma unroll
            for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(d2 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        } else {
          bf16* dst = static_cast<bf16*>(p.C) + (size_t)r * p.ldc + c0;
          if (p.e.accumulate) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 o = *reinterpret_cast<const uint4*>(dst + 8 * j);
              const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                v[8 * j + 2 * k] += __uint_as_float(w[k] << 16);
                v[8 * j + 2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
              }
            }
          }
          if (p.e.gate) {
            const uint4* gp = reinterpret_cast<const uint4*>(static_cast<const bf16*>(p.e.gate) + (size_t)r * p.e.gate_ld + c0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 o = __ldg(gp + j);
              const uint32_t w[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                v[8 * j + 2 * k] = __uint_as_float(w[k] << 16) > 0.f ? v[8 * j + 2 * k] : 0.f;
                v[8 * j + 2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u) > 0.f ? v[8 * j + 2 * k + 1] : 0.f;
              }
            }
          }
          uint32_t w[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            w[j] = *reinterpret_cast<uint32_t*>(&h);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(dst + 8 * j) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
          if (r2 >= 0) {
            bf16* d2 = static_cast<bf16*>(p.e.c2) + (size_t)r2 * p.e.c2_ld + c0;
#pragma unroll
            for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(d2 + 8 * j) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
          }
        }
      } else if (p.c_f32) {
        // ---- ragged edge / unaligned: generic per-element path
        float* C = static_cast<float*>(p.C);
        for (int j = 0; j < 32 && c0 + j < p.N; ++j) epi_store<float>(p.e, r, c0 + j, v[j], C, p.ldc);
      } else {
        bf16* C = static_cast<bf16*>(p.C);
        for (int j = 0; j < 32 && c0 + j < p.N; ++j) epi_store<bf16>(p.e, r, c0 + j, v[j], C, p.ldc);
      }
    }
    // this warp has finished reading the accumulator stage: hand it back to the MMA issuer
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tmem_empty + acc_stage);
    }  // units
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(f);
  });
  return fn;
}

// 2-D bf16 tensor map: inner extent `inner` (contiguous), outer extent `outer`, row pitch ld elements; box {64, box_rows}
int make_map(CUtensorMap* out, const void* base, long long inner, long long outer, long long ld, int box_rows) {
  typedef std::tuple<const void*, long long, long long, long long, int> Key;
  static std::map<Key, CUtensorMap> cache;
  static std::mutex mu;
  Key key(base, inner, outer, ld, box_rows);
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return REGAT_OK; }
  }
  EncodeFn fn = encode_fn();
  REGAT_REQUIRE(fn, REGAT_ERR_CUDA, "gemm_tc: cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  REGAT_REQUIRE(r == CUDA_SUCCESS, REGAT_ERR_CUDA, "gemm_tc: cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld ld=%lld", (int)r,
                inner, outer, ld);
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 4096) cache.clear();
  cache[key] = *out;
  return REGAT_OK;
}

template <int BN, int STAGES, int ACC>
int launch_cfg(bool a_mn, bool b_mn, const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, int ctas, cudaStream_t st) {
  constexpr size_t smem = (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + 1024 /*align slack*/ + (2 * STAGES + 2 * ACC) * 8 + 16;
#define REGAT_TC_CASE(AM, BMN)                                                                              \
  {                                                                                                         \
    auto kern = gemm_tc_kernel<BN, STAGES, ACC, AM, BMN>;                                                   \
    static bool attr_set = false;                                                                           \
    if (!attr_set) {                                                                                        \
      REGAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
      attr_set = true;                                                                                      \
    }                                                                                                       \
    kern<<<ctas, NTHREADS, smem, st>>>(ma, mb, p);                                                          \
  }
  if (!a_mn && b_mn) REGAT_TC_CASE(false, true)
  else if (!a_mn && !b_mn) REGAT_TC_CASE(false, false)
  else if (a_mn && b_mn) REGAT_TC_CASE(true, true)
  else REGAT_TC_CASE(true, false)
#undef REGAT_TC_CASE
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}

}  // namespace

bool gemm_tc_supported(int transA, int transB, int M, int N, int K, const void* A, int lda, const void* B, int ldb) {
  (void)transA; (void)transB; (void)M; (void)N; (void)K;
  return aligned16(A) && aligned16(B) && lda % 8 == 0 && ldb % 8 == 0;
}

int gemm_tc(int transA, int transB, int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C, int ldc,
            int c_dtype, const EpiArgs& e, int split_k, cudaStream_t st, int c_block_cols, const long long* c_block_off) {
  if (M <= 0 || N <= 0) return REGAT_OK;
  REGAT_REQUIRE(K > 0, REGAT_ERR_SHAPE, "gemm: K must be positive");
  REGAT_REQUIRE(gemm_tc_supported(transA, transB, M, N, K, A, lda, B, ldb), REGAT_ERR_ALIGN, "gemm_tc: unaligned operands");
  const bool a_mn = transA != 0;   // A stored [K, M]: M contiguous
  const bool b_mn = transB == 0;   // B stored [K, N]: N contiguous
  // tile width: 256 when there are enough column tiles to keep >= 1 wave busy, else 128
  static const int force_bn = [] { const char* s = getenv("REGAT_TC_BN"); return s ? atoi(s) : 0; }();
  const int tiles_m = ceil_div(M, BM);
  int bn = (force_bn == 128 || force_bn == 256) ? force_bn : ((N >= 256 && tiles_m * ceil_div(N, 256) >= num_sms()) ? 256 : 128);
  const int tiles_n = ceil_div(N, bn);
  const int total_kb = ceil_div(K, BK);
  // split-K only for plain fp32 targets (weight gradients: few output tiles, very long K): fill ~2 CTAs per SM
  int splits = 1;
  const bool plain = c_dtype == REGAT_F32 && !e.bias && !e.addend && !e.relu && !e.gate && !e.c2 && !e.accumulate;
  static const int auto_split = [] { const char* s = getenv("REGAT_TC_SPLITK"); return s ? atoi(s) : 1; }();
  if (plain) {
    if (split_k > 1) splits = split_k;
    else if (auto_split && tiles_m * tiles_n < num_sms()) splits = (2 * num_sms()) / (tiles_m * tiles_n);
    splits = std::max(1, std::min(splits, total_kb / 8));
  }
  const int kbps = ceil_div(total_kb, splits);
  splits = ceil_div(total_kb, kbps);

  CUtensorMap ma, mb;
  if (!a_mn) REGAT_TRY(make_map(&ma, A, K, M, lda, BM)); else REGAT_TRY(make_map(&ma, A, M, K, lda, BK));
  if (!b_mn) REGAT_TRY(make_map(&mb, B, K, N, ldb, bn)); else REGAT_TRY(make_map(&mb, B, N, K, ldb, BK));

  TcParams p;
  p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.c_f32 = c_dtype == REGAT_F32; p.k_blocks_per_split = kbps; p.total_k_blocks = total_kb;
  p.atomic_out = splits > 1; p.C = C; p.e = e;
  p.c_block_cols = 0;
  for (int j = 0; j < 8; ++j) p.c_block_off[j] = 0;
  if (c_block_cols > 0) {
    REGAT_REQUIRE(c_dtype == REGAT_F32 && plain && c_block_cols % 32 == 0 && N % c_block_cols == 0 && N / c_block_cols <= 8 && c_block_off,
                  REGAT_ERR_ARG, "gemm_tc: column-block scatter needs a plain fp32 output and <= 8 blocks of a multiple of 32 columns");
    p.c_block_cols = c_block_cols;
    for (int j = 0; j < N / c_block_cols; ++j) p.c_block_off[j] = c_block_off[j];
  }
  if (splits > 1) {   // partials are atomically added: zero the destination(s)
    const int nb = c_block_cols > 0 ? N / c_block_cols : 1, bw = c_block_cols > 0 ? c_block_cols : N;
    for (int j = 0; j < nb; ++j)
      REGAT_CUDA(cudaMemset2DAsync(static_cast<float*>(C) + p.c_block_off[j], (size_t)ldc * 4, 0, (size_t)bw * 4, (size_t)M, st));
  }
  p.tiles_m = tiles_m; p.tiles_n = tiles_n; p.splits = splits;
  const int units = tiles_m * tiles_n * splits;
  // BN=256: 4 stages x 48 KB + 2 x 256 TMEM columns, one CTA per SM.  BN=128: 3 stages x 32 KB + 2 x 128 columns, two per SM.
  // REGAT_SM_RESERVE=n leaves n SMs free of persistent GEMM CTAs so that a concurrent NCCL all-reduce can make progress
  static const int reserve = [] { const char* s = getenv("REGAT_SM_RESERVE"); return s ? std::max(0, atoi(s)) : 0; }();
  const int sms = std::max(1, num_sms() - reserve);
  if (bn == 256) return launch_cfg<256, 4, 2>(a_mn, b_mn, ma, mb, p, std::min(units, sms), st);
  return launch_cfg<128, 3, 2>(a_mn, b_mn, ma, mb, p, std::min(units, 2 * sms), st);
}

}  // namespace regat
