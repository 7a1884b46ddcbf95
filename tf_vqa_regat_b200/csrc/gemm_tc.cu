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
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2.. = EPIW (4 or 8) epilogue warps
// (warp w reads TMEM lanes 32*(w%4)..; with 8, the two warps of a lane quarter interleave the 32-column blocks).  Every mbarrier wait is bounded and traps instead of hanging.
// Epilogue data path: tcgen05.ld hands each thread one accumulator ROW, which is the wrong shape for global memory
// (a warp store would touch 32 different lines).  Each epilogue warp therefore transposes 32x32 fp32 blocks through a
// private 4 KB XOR-swizzled shared-memory patch: row-per-lane in, 8 lanes per 128-byte row out, so every global access
// of the fused epilogue (C, accumulate, gate, addend, second output, split-K reductions) is a run of full 32-byte
// sectors, and the next block's tcgen05.ld is in flight while the current one drains.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace regat {
namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr uint32_t SPIN_LIMIT = 1u << 27;
constexpr int EPI_PATCH_BYTES = 32 * 32 * 4;   // one 32x32 fp32 transpose patch per epilogue warp

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
#pragma unroll 1
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
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
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
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// the "+r" ties keep every use of the loaded registers behind the wait
__device__ __forceinline__ void tmem_ld32_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                 "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                 "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// C / gate / c2 arrive as generic pointers; they are global memory by contract, and saying so keeps the accesses off the
// generic path (which the compiler must order against the shared-memory patch)
__device__ __forceinline__ void stg_v4(void* p, float4 v) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
__device__ __forceinline__ void stg_v2(void* p, uint2 v) {
  asm volatile("st.global.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(v.x), "r"(v.y));
}
__device__ __forceinline__ float4 ldg_v4(const void* p) {
  float4 v;
  asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg_nc_v4(const void* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint2 ldg_nc_v2(const void* p) {
  uint2 v;
  asm volatile("ld.global.nc.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 lds_v4(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ uint2 ldg_v2(const void* p) {
  uint2 v;
  asm volatile("ld.global.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

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

// Optional second product riding in the same launch.  It shares N, ldc, the output type and the (empty) epilogue with the first
// and differs only in what is listed here; its units follow the first problem's.
struct TcSecond {
  int units;                      // tiles_m * tiles_n (no split-K); 0 = absent
  int M, tiles_m, total_k_blocks, accumulate;
  void* C;
};

struct TcParams {
  int M, N, K, ldc, c_f32, k_blocks_per_split, total_k_blocks, atomic_out, tiles_m, tiles_n, splits;
  int c_block_cols;               // > 0: column block j = c / c_block_cols of C lives at C + c_block_off[j] (fp32 outputs only)
  long long c_block_off[8];
  long long c_block_off_or;       // OR of all block offsets (alignment test)
  void* C;
  EpiArgs e;
  TcSecond s2;
};

// Generic per-element epilogue for ragged edges and unaligned tensors (4 consecutive columns of one row).
__device__ __noinline__ void epi_slow(const TcParams& p, void* Cout, int accumulate, int r, int c, float4 x) {
  const float xs[4] = {x.x, x.y, x.z, x.w};
  EpiArgs e = p.e;
  e.accumulate = accumulate;
  for (int j = 0; j < 4 && c + j < p.N; ++j) {
    if (p.atomic_out) {
      float y = xs[j];
      if (e.alpha) y *= e.alpha[e.alpha_cols ? (c + j) / e.alpha_cols : 0];
      int cj = c + j;
      long long off = 0;
      if (p.c_block_cols) { const int jb = cj / p.c_block_cols; off = p.c_block_off[jb]; cj -= jb * p.c_block_cols; }
      atomicAdd(static_cast<float*>(Cout) + off + (size_t)r * p.ldc + cj, y);
    } else if (p.c_f32) {
      epi_store<float>(e, r, c + j, xs[j], static_cast<float*>(Cout), p.ldc);
    } else {
      epi_store<bf16>(e, r, c + j, xs[j], static_cast<bf16*>(Cout), p.ldc);
    }
  }
}

// Persistent: CTA c processes work units c, c+grid, ... where a unit = (m-tile, n-tile, k-split), n fastest so that
// concurrently running CTAs share A rows in L2.  ACC accumulator stages in TMEM (ACC*BN <= 512 columns) let the epilogue
// of unit i overlap the MMAs of unit i+1.
template <int BN, int STAGES, int ACC, int EPIW, bool A_MN, bool B_MN, bool PAIR>
__global__ void __launch_bounds__(64 + 32 * EPIW, BN == 128 ? 2 : 0) gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA,
                                                           const __grid_constant__ CUtensorMap mapB, const __grid_constant__ TcParams p,
                                                           const __grid_constant__ CUtensorMap mapA2,
                                                           const __grid_constant__ CUtensorMap mapB2) {
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = ACC * BN;
  static_assert(TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two");
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* tiles = smem;            // SWIZZLE_128B tiles need 1024-byte alignment: the dynamic window provides it (checked)
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  float* stage = reinterpret_cast<float*>(tiles + STAGES * STAGE_BYTES);                 // EPIW epilogue warps x 4 KB
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES + EPIW * EPI_PATCH_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;       // [ACC]
  uint64_t* tmem_empty = tmem_full + ACC;         // [ACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + ACC);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = p.tiles_n, splits = p.splits;
  // units [0, units1) belong to the launch's first problem, [units1, units) to the optional second one (host: longer K first)
  const int units1 = p.tiles_m * tiles_n * splits;
  const int units = units1 + (PAIR ? p.s2.units : 0);

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int a = 0; a < ACC; ++a) { mbar_init(tmem_full + a, 1); mbar_init(tmem_empty + a, EPIW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // unit -> (m0, n0, kb0, kb1)
  auto decode = [&](int u, int& m0, int& n0, int& kb0, int& kb1) {
    if (PAIR && u >= units1) {    // second problem: no split-K
      const int tile = u - units1;
      const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
      m0 = tm * BM; n0 = tn * BN; kb0 = 0; kb1 = p.s2.total_k_blocks;
      return;
    }
    const int tile = u / splits, z = u - tile * splits;
    const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
    m0 = tm * BM; n0 = tn * BN;
    kb0 = z * p.k_blocks_per_split;
    kb1 = min(p.total_k_blocks, kb0 + p.k_blocks_per_split);
  };

  if (warp == 0) {
    // ===== TMA producer: warp-uniform loop, one elected lane issues =====
    {
      const bool leader = elect_one();
      int s = 0;
      uint32_t ph = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int m0, n0, kb0, kb1;
        decode(u, m0, n0, kb0, kb1);
        const CUtensorMap* ma = (PAIR && u >= units1) ? &mapA2 : &mapA;
        const CUtensorMap* mb = (PAIR && u >= units1) ? &mapB2 : &mapB;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar + s, ph ^ 1);
          if (leader) {
          mbar_expect_tx(full_bar + s, STAGE_BYTES);
          unsigned char* sa = tiles + s * STAGE_BYTES;
          unsigned char* sb = sa + A_BYTES;
          if (!A_MN) {
            tma_load_2d(sa, ma, full_bar + s, kb * BK, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * (64 * BK * 2), ma, full_bar + s, m0 + c * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(sb, mb, full_bar + s, kb * BK, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c) tma_load_2d(sb + c * (64 * BK * 2), mb, full_bar + s, n0 + c * 64, kb * BK);
          }
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp walks the pipeline (uniform control flow), one elected lane issues =====
    // The issue loop is on the critical path of every k-block (4 MMAs of 128 x BN x 16 take only ~128 BN/256 ns), so it is
    // kept to a wait, four descriptor adds and the instructions themselves: descriptors are built once and advanced by
    // adding the byte offset >> 4 to their start-address field (shared memory addresses fit its 14 bits without carry).
    constexpr uint32_t idesc = make_idesc(BN, A_MN, B_MN);
    constexpr uint32_t KSTEP_A = A_MN ? 2048u : 32u, KSTEP_B = B_MN ? 2048u : 32u;   // bytes per UMMA_K step
    const uint32_t tiles_addr = smem_u32(tiles);
    const uint64_t da0 = A_MN ? make_desc(tiles_addr, 64 * BK * 2, 1024) : make_desc(tiles_addr, 16, 1024);
    const uint64_t db0 = B_MN ? make_desc(tiles_addr + A_BYTES, 64 * BK * 2, 1024) : make_desc(tiles_addr + A_BYTES, 16, 1024);
    const bool leader = elect_one();
    int s = 0, ui = 0;
    uint32_t ph = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++ui) {
      int m0, n0, kb0, kb1;
      decode(u, m0, n0, kb0, kb1);
      const int a = ui % ACC;
      const uint32_t aph = (ui / ACC) & 1;
      mbar_wait(tmem_empty + a, aph ^ 1);          // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(a * BN);
      uint32_t acc_flag = 0u;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar + s, ph);
        tc_fence_after();
        if (leader) {
          const uint64_t da = da0 + (uint64_t)((uint32_t)s * (STAGE_BYTES >> 4));
          const uint64_t db = db0 + (uint64_t)((uint32_t)s * (STAGE_BYTES >> 4));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            umma_bf16(tmem_d, da + (uint64_t)(k * (KSTEP_A >> 4)), db + (uint64_t)(k * (KSTEP_B >> 4)), idesc, k == 0 ? acc_flag : 1u);
          }
          umma_commit(empty_bar + s);          // frees the smem slot once these MMAs have read it
        }
        acc_flag = 1u;
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      if (leader) umma_commit(tmem_full + a);  // accumulator of this unit complete
      __syncwarp();
    }
  } else {
    // ===== epilogue: TMEM -> registers -> swizzled smem patch -> fused epilogue -> coalesced global =====
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    float* patch = stage + (warp - 2) * (32 * 32);          // this warp's 32 x 32 fp32 transpose patch
    const int lr = lane >> 3, lc = lane & 7; // drain phase: lane handles the 16-byte column group lc of rows lr + 4 i
    const size_t esz = p.c_f32 ? 4 : 2;
    // vector path needs 16-byte aligned rows for every tensor touched with vector accesses
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(p.C) | (PAIR ? reinterpret_cast<uintptr_t>(p.s2.C) : 0) | ((size_t)p.ldc * esz)) & 15) == 0 &&
                        (!p.e.bias || (reinterpret_cast<uintptr_t>(p.e.bias) & 15) == 0) &&
                        (!p.e.addend || ((reinterpret_cast<uintptr_t>(p.e.addend) | ((size_t)p.e.addend_ld * 4)) & 15) == 0) &&
                        (!p.e.gate || ((reinterpret_cast<uintptr_t>(p.e.gate) | ((size_t)p.e.gate_ld * esz)) & 15) == 0) &&
                        (!p.e.c2 || ((reinterpret_cast<uintptr_t>(p.e.c2) | ((size_t)p.e.c2_ld * esz)) & 15) == 0) &&
                        (p.e.alpha_cols % 32 == 0) && (p.c_block_off_or & 3) == 0;
    const float alpha0 = (p.e.alpha && p.e.alpha_cols == 0) ? __ldg(p.e.alpha) : 1.f;
    int ui = 0;
    for (int u = blockIdx.x; u < units; u += gridDim.x, ++ui) {
      int m0, n0, kb0, kb1;
      decode(u, m0, n0, kb0, kb1);
      // the few things in which the launch's second problem differs from the first
      const bool sec = PAIR && u >= units1;
      const int uM = sec ? p.s2.M : p.M;
      void* const uC = sec ? p.s2.C : p.C;
      const int uAcc = sec ? p.s2.accumulate : p.e.accumulate;
      const int acc_stage = ui % ACC;
      // read-modify-write epilogues (bf16 C): pull the old C / gate tiles into L2 while the MMAs of this unit are still running
      if (!p.c_f32 && (uAcc || p.e.gate)) {
        const int et = (warp - 2) * 32 + lane;
        constexpr int LPR = BN / 64;                    // 128-byte lines per tile row
        for (int id = et; id < BM * LPR; id += 32 * EPIW) {
          const int row = m0 + id / LPR, col = n0 + (id % LPR) * 64;
          if (row < uM && col < p.N) {
            if (uAcc) asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const bf16*>(uC) + (size_t)row * p.ldc + col));
            if (p.e.gate) asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const bf16*>(p.e.gate) + (size_t)row * p.e.gate_ld + col));
          }
        }
      }
      mbar_wait(tmem_full + acc_stage, (ui / ACC) & 1);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(acc_stage * BN) + ((uint32_t)(q * 32) << 16);
      const int rbase = m0 + q * 32;
      // 32-column blocks this warp really has to move (none if its rows or the unit's K range are empty)
      const int nblk = (rbase < uM && kb1 > kb0) ? min(BN / 32, (p.N - n0 + 31) / 32) : 0;
      // per-row terms of the 8 rows this lane drains, hoisted out of the column loop
      float rs[8];
      int arow[8], r2[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = min(rbase + lr + 4 * i, uM - 1);
        rs[i] = (p.e.addend && p.e.row_scale) ? __ldg(p.e.row_scale + r) : 1.f;
        arow[i] = p.e.addend ? (r / p.e.addend_rows) * p.e.addend_ld : 0;
        r2[i] = -1;
        if (p.e.c2) {
          const int rr = r % p.e.c2_rows_in;
          if (rr < p.e.c2_rows_keep) r2[i] = (r / p.e.c2_rows_in) * p.e.c2_rows_keep + rr;
        }
      }
      uint32_t acc[32];
      constexpr int CSTEP = EPIW / 4;                      // warps sharing a lane quarter interleave the column blocks
      const int cfirst = (warp - 2) >> 2;
      if (cfirst < nblk) tmem_ld32_issue(tmem_acc + (uint32_t)(cfirst * 32), acc);
#pragma unroll 1
      for (int cc = cfirst; cc < nblk; cc += CSTEP) {
        const int c = n0 + cc * 32 + lc * 4;               // first of this lane's 4 columns in the drain phase
        const bool vec = vec_ok && c + 4 <= p.N;
        // column terms: issued before the TMEM wait so their latency is hidden
        float al = alpha0;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vec) {
          if (p.e.alpha && p.e.alpha_cols) al = __ldg(p.e.alpha + c / p.e.alpha_cols);
          if (p.e.bias) bv = __ldg(reinterpret_cast<const float4*>(p.e.bias + c));
        }
        // bf16 outputs: the read-modify-write terms (old C for accumulate, the relu gate) are requested here too -- their
        // DRAM latency then runs under the TMEM wait and the transpose instead of stalling the drain
        const int rcl = min(rbase + lr, uM - 1);      // clamped first row of this lane (loads only)
        uint2 oldc[8], gt[8];
        const bool pre_acc = vec && !p.c_f32 && uAcc, pre_gate = vec && !p.c_f32 && p.e.gate;
        if (pre_acc) {
          const bf16* c0p = static_cast<const bf16*>(uC) + c;
#pragma unroll
          for (int i = 0; i < 8; ++i) oldc[i] = ldg_v2(c0p + (size_t)min(rcl + 4 * i, uM - 1) * p.ldc);
        }
        if (pre_gate) {
          const bf16* g0 = static_cast<const bf16*>(p.e.gate) + c;
#pragma unroll
          for (int i = 0; i < 8; ++i) gt[i] = ldg_nc_v2(g0 + (size_t)min(rcl + 4 * i, uM - 1) * p.e.gate_ld);
        }
        tmem_ld32_wait(acc);
        // transpose in: lane = row, 16-byte group g lands at slot g ^ (row & 7)  (conflict-free both ways)
        {
          float4* prow = reinterpret_cast<float4*>(patch + lane * 32);
#pragma unroll
          for (int g = 0; g < 8; ++g)
            prow[g ^ (lane & 7)] = make_float4(__uint_as_float(acc[4 * g]), __uint_as_float(acc[4 * g + 1]),
                                               __uint_as_float(acc[4 * g + 2]), __uint_as_float(acc[4 * g + 3]));
        }
        if (cc + CSTEP < nblk) {
          tmem_ld32_issue(tmem_acc + (uint32_t)((cc + CSTEP) * 32), acc);   // overlaps the drain below
        } else {
          // every TMEM read of this unit has completed: hand the accumulator stage back to the MMA issuer early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty + acc_stage);
        }
        __syncwarp();
        // column-block scatter (weight gradients of side-by-side layers): block j of C lives at C + c_block_off[j]
        long long cboff = 0;
        int cl = c;
        if (p.c_block_cols) { const int jb = c / p.c_block_cols; cboff = p.c_block_off[jb]; cl = c - jb * p.c_block_cols; }
        if (!vec) {
          // ragged edge / unaligned tensors: generic per-element path (kept out of line: it must not bloat the hot loop)
          if (c < p.N) {
#pragma unroll 1
            for (int i = 0; i < 8; ++i) {
              const int rl = lr + 4 * i, r = rbase + rl;
              if (r < uM) epi_slow(p, uC, uAcc, r, c, reinterpret_cast<const float4*>(patch + rl * 32)[lc ^ (rl & 7)]);
            }
          }
        } else {
          // Staged, branch-free drain: every load of a stage is issued before the first use (shared-memory and global
          // latencies are paid once per stage, not once per row); rows past M are clamped for loads, predicated for stores.
          float4 x[8];
          const uint32_t pbase = smem_u32(patch);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rl = lr + 4 * i;
            x[i] = lds_v4(pbase + (uint32_t)(rl * 128 + ((lc ^ (rl & 7)) << 4)));
          }
          if (p.e.addend) {
            float4 a[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = ldg_nc_v4(p.e.addend + arow[i] + c);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              x[i].x = fmaf(rs[i], a[i].x, x[i].x); x[i].y = fmaf(rs[i], a[i].y, x[i].y);
              x[i].z = fmaf(rs[i], a[i].z, x[i].z); x[i].w = fmaf(rs[i], a[i].w, x[i].w);
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            x[i].x = fmaf(x[i].x, al, bv.x); x[i].y = fmaf(x[i].y, al, bv.y); x[i].z = fmaf(x[i].z, al, bv.z); x[i].w = fmaf(x[i].w, al, bv.w);
          }
          if (p.e.relu) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              x[i].x = fmaxf(x[i].x, 0.f); x[i].y = fmaxf(x[i].y, 0.f); x[i].z = fmaxf(x[i].z, 0.f); x[i].w = fmaxf(x[i].w, 0.f);
            }
          }
          const int rows_ok = uM - rbase - lr;        // row i of this lane is valid iff 4 i < rows_ok
          if (p.c_f32) {
            float* dst0 = static_cast<float*>(uC) + cboff + cl;
            if (p.atomic_out) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (4 * i < rows_ok) red_add_v4(dst0 + (size_t)(rbase + lr + 4 * i) * p.ldc, x[i]);
            } else {
              if (uAcc) {
                float4 o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = ldg_v4(dst0 + (size_t)min(rcl + 4 * i, uM - 1) * p.ldc);
#pragma unroll
                for (int i = 0; i < 8; ++i) { x[i].x += o[i].x; x[i].y += o[i].y; x[i].z += o[i].z; x[i].w += o[i].w; }
              }
              if (p.e.gate) {
                const float* g0 = static_cast<const float*>(p.e.gate) + c;
                float4 g[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) g[i] = ldg_nc_v4(g0 + (size_t)min(rcl + 4 * i, uM - 1) * p.e.gate_ld);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  x[i].x = g[i].x > 0.f ? x[i].x : 0.f; x[i].y = g[i].y > 0.f ? x[i].y : 0.f;
                  x[i].z = g[i].z > 0.f ? x[i].z : 0.f; x[i].w = g[i].w > 0.f ? x[i].w : 0.f;
                }
              }
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (4 * i < rows_ok) stg_v4(dst0 + (size_t)(rbase + lr + 4 * i) * p.ldc, x[i]);
              if (p.e.c2) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  if (4 * i < rows_ok && r2[i] >= 0) stg_v4(static_cast<float*>(p.e.c2) + (size_t)r2[i] * p.e.c2_ld + c, x[i]);
              }
            }
          } else {
            bf16* dst0 = static_cast<bf16*>(uC) + c;
            if (uAcc) {
#pragma unroll
              for (int i = 0; i < 8; ++i) { x[i].x += bf_lo(oldc[i].x); x[i].y += bf_hi(oldc[i].x); x[i].z += bf_lo(oldc[i].y); x[i].w += bf_hi(oldc[i].y); }
            }
            if (p.e.gate) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                x[i].x = bf_lo(gt[i].x) > 0.f ? x[i].x : 0.f; x[i].y = bf_hi(gt[i].x) > 0.f ? x[i].y : 0.f;
                x[i].z = bf_lo(gt[i].y) > 0.f ? x[i].z : 0.f; x[i].w = bf_hi(gt[i].y) > 0.f ? x[i].w : 0.f;
              }
            }
            uint2 w[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) w[i] = make_uint2(pack_bf16(x[i].x, x[i].y), pack_bf16(x[i].z, x[i].w));
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (4 * i < rows_ok) stg_v2(dst0 + (size_t)(rbase + lr + 4 * i) * p.ldc, w[i]);
            if (p.e.c2) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (4 * i < rows_ok && r2[i] >= 0) stg_v2(static_cast<bf16*>(p.e.c2) + (size_t)r2[i] * p.e.c2_ld + c, w[i]);
            }
          }
        }
        __syncwarp();    // the patch is rewritten by the next block
      }
      if (cfirst >= nblk) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty + acc_stage);
      }
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

template <int BN, int STAGES, int ACC, int EPIW>
int launch_cfg(bool a_mn, bool b_mn, const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, const CUtensorMap& ma2,
               const CUtensorMap& mb2, int ctas, cudaStream_t st) {
  const bool pair = p.s2.units > 0;
  constexpr size_t smem = (size_t)STAGES * (BM * BK * 2 + BN * BK * 2) + EPIW * EPI_PATCH_BYTES + (2 * STAGES + 2 * ACC) * 8 + 16;
  // the shared-memory attribute is per (function, device): one flag per device, set under a lock
#define REGAT_TC_CASE(AM, BMN)                                                                              \
  if (pair) REGAT_TC_CASE2(AM, BMN, true) else REGAT_TC_CASE2(AM, BMN, false)
#define REGAT_TC_CASE2(AM, BMN, PR)                                                                         \
  {                                                                                                         \
    auto kern = gemm_tc_kernel<BN, STAGES, ACC, EPIW, AM, BMN, PR>;                                               \
    static std::mutex mu;                                                                                   \
    static bool attr_set[64] = {};                                                                          \
    int dev = 0;                                                                                            \
    REGAT_CUDA(cudaGetDevice(&dev));                                                                        \
    {                                                                                                       \
      std::lock_guard<std::mutex> lk(mu);                                                                   \
      if (dev < 0 || dev >= 64 || !attr_set[dev]) {                                                         \
        REGAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
        if (dev >= 0 && dev < 64) attr_set[dev] = true;                                                     \
      }                                                                                                     \
    }                                                                                                       \
    kern<<<ctas, 64 + 32 * EPIW, smem, st>>>(ma, mb, p, ma2, mb2);                                                \
  }
  if (!a_mn && b_mn) REGAT_TC_CASE(false, true)
  else if (!a_mn && !b_mn) REGAT_TC_CASE(false, false)
  else if (a_mn && b_mn) REGAT_TC_CASE(true, true)
  else REGAT_TC_CASE(true, false)
#undef REGAT_TC_CASE
#undef REGAT_TC_CASE2
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}

struct Prepared {
  CUtensorMap ma, mb;
  TcParams p;
  bool a_mn, b_mn;
  int bn;
};

// tile width, split-K and tensor maps of one problem; zeroes split-K destinations on `st`
int prepare(int transA, int transB, int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C, int ldc,
            int c_dtype, const EpiArgs& e, int split_k, cudaStream_t st, int c_block_cols, const long long* c_block_off, int force_bn_arg,
            Prepared& out) {
  REGAT_REQUIRE(K > 0, REGAT_ERR_SHAPE, "gemm: K must be positive");
  REGAT_REQUIRE(gemm_tc_supported(transA, transB, M, N, K, A, lda, B, ldb), REGAT_ERR_ALIGN, "gemm_tc: unaligned operands");
  const bool a_mn = transA != 0;   // A stored [K, M]: M contiguous
  const bool b_mn = transB == 0;   // B stored [K, N]: N contiguous
  // tile width: 256 when there are enough column tiles to keep >= 1 wave busy, else 128
  static const int force_bn_env = [] { const char* s = getenv("REGAT_TC_BN"); return s ? atoi(s) : 0; }();
  const int force_bn = force_bn_arg ? force_bn_arg : force_bn_env;
  const int tiles_m = ceil_div(M, BM);
  int bn = (force_bn == 64 || force_bn == 128 || force_bn == 256) ? force_bn : ((N >= 256 && tiles_m * ceil_div(N, 256) >= num_sms()) ? 256 : 128);
  // skinny problems (a batch-sized M, long K): the 128-wide tiling leaves most SMs idle behind a serial K loop that is
  // bound by TMA latency at 3 stages -- narrower tiles double the CTAs and the 24 KB stages allow an 8-deep ring
  static const int skinny = [] { const char* s = getenv("REGAT_TC_SKINNY"); return s ? atoi(s) : 1; }();
  if (!force_bn && skinny && bn == 128 && tiles_m <= 2 && tiles_m * ceil_div(N, 128) * 2 <= num_sms() && K >= 512) bn = 64;
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

  if (!a_mn) REGAT_TRY(make_map(&out.ma, A, K, M, lda, BM)); else REGAT_TRY(make_map(&out.ma, A, M, K, lda, BK));
  if (!b_mn) REGAT_TRY(make_map(&out.mb, B, K, N, ldb, bn)); else REGAT_TRY(make_map(&out.mb, B, N, K, ldb, BK));

  TcParams& p = out.p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K; p.ldc = ldc; p.c_f32 = c_dtype == REGAT_F32; p.k_blocks_per_split = kbps; p.total_k_blocks = total_kb;
  p.atomic_out = splits > 1; p.C = C; p.e = e;
  p.c_block_cols = 0; p.c_block_off_or = 0;
  for (int j = 0; j < 8; ++j) p.c_block_off[j] = 0;
  if (c_block_cols > 0) {
    REGAT_REQUIRE(c_dtype == REGAT_F32 && plain && c_block_cols % 32 == 0 && N % c_block_cols == 0 && N / c_block_cols <= 8 && c_block_off,
                  REGAT_ERR_ARG, "gemm_tc: column-block scatter needs a plain fp32 output and <= 8 blocks of a multiple of 32 columns");
    p.c_block_cols = c_block_cols;
    for (int j = 0; j < N / c_block_cols; ++j) { p.c_block_off[j] = c_block_off[j]; p.c_block_off_or |= c_block_off[j]; }
  }
  if (splits > 1) {   // partials are atomically added: zero the destination(s)
    const int nb = c_block_cols > 0 ? N / c_block_cols : 1, bw = c_block_cols > 0 ? c_block_cols : N;
    for (int j = 0; j < nb; ++j)
      REGAT_CUDA(cudaMemset2DAsync(static_cast<float*>(C) + p.c_block_off[j], (size_t)ldc * 4, 0, (size_t)bw * 4, (size_t)M, st));
  }
  p.tiles_m = tiles_m; p.tiles_n = tiles_n; p.splits = splits;
  out.a_mn = a_mn; out.b_mn = b_mn; out.bn = bn;
  return REGAT_OK;
}

int launch_prepared(const Prepared& x, const Prepared* y, cudaStream_t st) {
  TcParams p = x.p;
  if (y) {
    p.s2.units = y->p.tiles_m * y->p.tiles_n; p.s2.M = y->p.M; p.s2.tiles_m = y->p.tiles_m; p.s2.total_k_blocks = y->p.total_k_blocks;
    p.s2.accumulate = y->p.e.accumulate; p.s2.C = y->p.C;
  }
  const int units = p.tiles_m * p.tiles_n * p.splits + p.s2.units;
  // BN=256: 4 stages x 48 KB + 2 x 256 TMEM columns, one CTA per SM.  BN=128: 3 stages x 32 KB + 2 x 128 columns, two per SM.
  // REGAT_SM_RESERVE=n leaves n SMs free of persistent GEMM CTAs so that a concurrent NCCL all-reduce can make progress
  static const int reserve = [] { const char* s = getenv("REGAT_SM_RESERVE"); return s ? std::max(0, atoi(s)) : 0; }();
  const int sms = std::max(1, num_sms() - reserve);
  const CUtensorMap& ma2 = y ? y->ma : x.ma;
  const CUtensorMap& mb2 = y ? y->mb : x.mb;
  if (x.bn == 256) return launch_cfg<256, 4, 2, 8>(x.a_mn, x.b_mn, x.ma, x.mb, p, ma2, mb2, std::min(units, sms), st);
  if (x.bn == 64) return launch_cfg<64, 8, 2, 4>(x.a_mn, x.b_mn, x.ma, x.mb, p, ma2, mb2, std::min(units, sms), st);
  return launch_cfg<128, 3, 2, 4>(x.a_mn, x.b_mn, x.ma, x.mb, p, ma2, mb2, std::min(units, 2 * sms), st);
}

}  // namespace

bool gemm_tc_supported(int transA, int transB, int M, int N, int K, const void* A, int lda, const void* B, int ldb) {
  (void)transA; (void)transB; (void)M; (void)N; (void)K;
  return aligned16(A) && aligned16(B) && lda % 8 == 0 && ldb % 8 == 0;
}

int gemm_tc(int transA, int transB, int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C, int ldc,
            int c_dtype, const EpiArgs& e, int split_k, cudaStream_t st, int c_block_cols, const long long* c_block_off) {
  if (M <= 0 || N <= 0) return REGAT_OK;
  Prepared x;
  REGAT_TRY(prepare(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, c_dtype, e, split_k, st, c_block_cols, c_block_off, 0, x));
  return launch_prepared(x, nullptr, st);
}

// Two independent products in ONE launch: the persistent CTAs walk the units of c0, then those of c1, so a part-filled last
// wave of one problem is topped up with tiles of the other.  Pass the problem with the longer K first.  The second problem may
// differ from the first in M, K, its operands, its output pointer and the accumulate flag only (same N, ldc, output type, no
// other epilogue term, no split-K); anything else falls back to two launches.
int gemm_tc_pair(const GemmCall& c0, const GemmCall& c1, cudaStream_t st) {
  auto single = [&](const GemmCall& c) {
    return gemm_tc(c.transA, c.transB, c.M, c.N, c.K, c.A, c.lda, c.B, c.ldb, c.C, c.ldc, c.c_dtype, c.e, 1, st);
  };
  if (c0.M <= 0 || c0.N <= 0) return single(c1);
  if (c1.M <= 0 || c1.N <= 0) return single(c0);
  auto bare = [](EpiArgs e) { e.accumulate = 0; EpiArgs z; memset(&z, 0, sizeof(z)); return memcmp(&e, &z, sizeof(z)) == 0; };
  static const int no_pair = [] { const char* s = getenv("REGAT_TC_PAIR"); return s ? atoi(s) == 0 : 0; }();
  bool ok = !no_pair && c0.N == c1.N && c0.ldc == c1.ldc && c0.c_dtype == c1.c_dtype && c0.transA == c1.transA && c0.transB == c1.transB &&
            bare(c0.e) && bare(c1.e) && c0.N % 8 == 0 && aligned16(c0.C) && aligned16(c1.C) && c0.ldc % 8 == 0;
  Prepared x, y;
  if (ok) {
    REGAT_TRY(prepare(c0.transA, c0.transB, c0.M, c0.N, c0.K, c0.A, c0.lda, c0.B, c0.ldb, c0.C, c0.ldc, c0.c_dtype, c0.e, 1, st, 0, nullptr, 0, x));
    REGAT_TRY(prepare(c1.transA, c1.transB, c1.M, c1.N, c1.K, c1.A, c1.lda, c1.B, c1.ldb, c1.C, c1.ldc, c1.c_dtype, c1.e, 1, st, 0, nullptr, x.bn, y));
    ok = x.p.splits == 1 && y.p.splits == 1 && x.bn == y.bn && x.p.tiles_n == y.p.tiles_n;
  }
  if (!ok) { REGAT_TRY(single(c0)); return single(c1); }
  return launch_prepared(x, &y, st);
}

}  // namespace regat
