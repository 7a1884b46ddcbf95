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
// Epilogue data path: tcgen05.ld (32x32b.x32) hands each thread 32 consecutive columns of one accumulator ROW -- 64 bytes
// of a bf16 output, 128 of an fp32 one, i.e. whole 32-byte sectors.  Each thread finishes its piece in registers and
// stores it with 256-bit accesses; the global-memory terms of the fused epilogue (old C, relu gate, addend) are requested
// before the TMEM wait, and the next block's tcgen05.ld is in flight while the current one is processed.  (The first
// version transposed 32x32 blocks through shared memory to get row-contiguous warp stores; measured on the step's shapes it
// cost ~6 us per 128x256 unit -- latency of four dependent phases per block with two warps per scheduler -- which bounded
// every K <= 1024 product and left a ~6 us tail on the others.)
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
// ---- CTA pair (cta_group::2): two CTAs of a cluster on one TPC share one 256-row MMA; CTA rank 0 issues it
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// each CTA of the pair loads its own half of the operands; the bytes are counted on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// completion of the pair's MMAs arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
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
// generic path
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
__device__ __forceinline__ uint4 ldg_v4u(const void* p) {
  uint4 v;
  asm volatile("ld.global.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ldg_nc_v4u(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_v4u(void* p, const uint32_t* w) {
  asm volatile("st.global.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]));
}
// 256-bit stores (sm_100): one full 32-byte sector per thread and instruction
__device__ __forceinline__ void stg_v8u(void* p, const uint32_t* w) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]),
               "r"(w[5]), "r"(w[6]), "r"(w[7]));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
// packed fp32x2 arithmetic (sm_100): (a0, a1) += (b0, b1);  (a0, a1) = (a0, a1) * s + (b0, b1)
__device__ __forceinline__ void fadd2(uint32_t& a0, uint32_t& a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%2, %3};\n\tadd.rn.f32x2 ra, ra, rb;\n\tmov.b64 {%0, %1}, ra;\n\t}"
      : "+r"(a0), "+r"(a1)
      : "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
__device__ __forceinline__ void ffma2(uint32_t& a0, uint32_t& a1, float s, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rs;\n\tmov.b64 ra, {%0, %1};\n\tmov.b64 rb, {%3, %4};\n\tmov.b64 rs, {%2, %2};\n\tfma.rn.f32x2 ra, ra, rs, rb;\n\tmov.b64 {%0, %1}, ra;\n\t}"
      : "+r"(a0), "+r"(a1)
      : "r"(__float_as_uint(s)), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
// relu folded into the conversion
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
  uint32_t w;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(hi), "f"(lo));
  return w;
}
// 0xffff in each half whose bf16 value is > 0
__device__ __forceinline__ uint32_t bf16x2_gt0_mask(uint32_t g) {
  return __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&g), __float2bfloat162_rn(0.f));
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
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn, int m = BM) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
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
  int dbg;                        // experiments (REGAT_TC_DBG): 1 = no stores, 2 = no TMEM loads -- results are wrong
  long long* trace;               // measurement aid (regat_gemm_trace): per-unit clock64 stamps of CTA 0's warp roles, or null
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
// CTA2: the launch is made of clusters of two CTAs (one TPC).  A unit is a 256 x BN tile: CTA rank r of the pair loads rows
// m0 + 128 r of A and columns n0 + (BN/2) r of B into its own shared memory, the leader (rank 0) issues ONE
// tcgen05.mma.cta_group::2 of 256 x BN x 16 per k-step that reads both CTAs' shared memory, and each CTA drains the 128
// accumulator rows that live in its own TMEM.  Per SM that halves the B traffic (TMA writes and MMA reads of shared memory),
// which is what bounds the single-CTA 128 x 256 tile.
// EPI selects how much of the fused epilogue is compiled in.  The epilogue is a long chain of optional terms; with all of them
// in one loop body the executed path hops through ~60 KB of code, the loop no longer fits the instruction cache next to the
// producer / MMA loops, and the epilogue warps run at ~4 cycles per instruction (measured: 4.5 us per 128 x 256 unit, more
// than the unit's K = 1024 main loop).  The host picks the smallest kind that covers the call (epi_kind()):
//   0 generic (everything, ragged / unaligned tensors included)      2 bf16 output with read-modify-write terms (old C, gate,
//   1 bf16 output, optional bias and relu                               addend, second output)
//   3 fp32 output, optional bias; plain store, split-K red or column-block scatter
template <int BN, int STAGES, int ACC, int EPIW, bool A_MN, bool B_MN, bool PAIR, bool CTA2, int EPI>
__global__ void __launch_bounds__(64 + 32 * EPIW, BN == 128 ? 2 : 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA,
                                                           const __grid_constant__ CUtensorMap mapB, const __grid_constant__ TcParams p,
                                                           const __grid_constant__ CUtensorMap mapA2,
                                                           const __grid_constant__ CUtensorMap mapB2) {
  constexpr int BNL = CTA2 ? BN / 2 : BN;         // B columns this CTA loads
  constexpr int TM = CTA2 ? 2 * BM : BM;          // rows of a unit
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BNL * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = ACC * BN;
  static_assert(TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM allocation must be a power of two");
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* tiles = smem;            // SWIZZLE_128B tiles need 1024-byte alignment: the dynamic window provides it (checked)
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;       // [ACC]
  uint64_t* tmem_empty = tmem_full + ACC;         // [ACC]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + ACC);
  float* bias_stage = reinterpret_cast<float*>(tmem_slot + 4);       // EPIW x 128 floats: each epilogue warp's bias columns of a unit

  // programmatic dependent launch: let the next kernel of the stream be scheduled as soon as resources free up (it blocks in its
  // own griddepcontrol.wait until this grid has completed)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = CTA2 ? cluster_ctarank() : 0u;            // rank in the CTA pair
  const int worker = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int workers = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int tiles_n = p.tiles_n, splits = p.splits;
  // units [0, units1) belong to the launch's first problem, [units1, units) to the optional second one (host: longer K first)
  const int units1 = p.tiles_m * tiles_n * splits;
  const int units = units1 + (PAIR ? p.s2.units : 0);

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    // CTA pair: the epilogue warps of BOTH CTAs hand an accumulator stage back to the leader
    for (int a = 0; a < ACC; ++a) { mbar_init(tmem_full + a, 1); mbar_init(tmem_empty + a, CTA2 ? 2 * EPIW : EPIW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) { if (CTA2) tmem_alloc_pair(tmem_slot, TMEM_COLS); else tmem_alloc(tmem_slot, TMEM_COLS); }
  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();     // pair: the peer's barriers exist before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (barriers, TMEM, tensor-map prefetch) touched nothing the previous kernel of the stream produces; from here
  // on it does (no-op unless launched with the programmatic attribute)
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // unit -> (m0, n0, kb0, kb1)
  auto decode = [&](int u, int& m0, int& n0, int& kb0, int& kb1) {
    if (PAIR && u >= units1) {    // second problem: no split-K
      const int tile = u - units1;
      const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
      m0 = tm * TM; n0 = tn * BN; kb0 = 0; kb1 = p.s2.total_k_blocks;
      return;
    }
    const int tile = u / splits, z = u - tile * splits;
    const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
    m0 = tm * TM; n0 = tn * BN;
    kb0 = z * p.k_blocks_per_split;
    kb1 = min(p.total_k_blocks, kb0 + p.k_blocks_per_split);
  };

  // measurement aid: CTA 0 stamps clock64 at the phase boundaries of each unit; slot layout trace[8 + 8 * unit + k]
  long long* const tr = (p.trace && blockIdx.x == 0) ? p.trace : nullptr;
  auto stamp = [&](int ui, int k) { if (tr && ui < 14) tr[8 + 8 * ui + k] = clock64(); };
  if (tr && threadIdx.x == 0) tr[0] = clock64();
  // Unit of this worker in round `it` (-1: none left).  Two-problem launches mix long-K and short-K units (long first): complete
  // odd rounds run in reverse worker order, so the workers that drew a second long unit are not the ones that get the extra
  // unit of the last, partial round (dQ / dKV input gradients: 192 -> 160 k-blocks on the most loaded worker).
  auto unit_at = [&](int it) -> int {
    const int base = it * workers;
    if (base >= units) return -1;
    if (PAIR && (it & 1) && base + workers <= units) return base + workers - 1 - worker;
    const int u = base + worker;
    return u < units ? u : -1;
  };
  if (warp == 0) {
    // ===== TMA producer: warp-uniform loop, one elected lane issues =====
    {
      const bool leader = elect_one();
      int s = 0;
      uint32_t ph = 0;
      for (int pui = 0, u; (u = unit_at(pui)) >= 0; ++pui) {
        int m0, n0, kb0, kb1;
        decode(u, m0, n0, kb0, kb1);
        const CUtensorMap* ma = (PAIR && u >= units1) ? &mapA2 : &mapA;
        const CUtensorMap* mb = (PAIR && u >= units1) ? &mapB2 : &mapB;
        const int ma0 = m0 + (int)crank * BM;          // this CTA's rows of A / columns of B
        const int nb0 = n0 + (int)crank * BNL;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar + s, ph ^ 1);            // the pair's MMAs have released this CTA's slot
          if (leader) {
          if (kb == kb0) stamp(pui, 0);
          if (kb == kb1 - 1) stamp(pui, 1);
          unsigned char* sa = tiles + s * STAGE_BYTES;
          unsigned char* sb = sa + A_BYTES;
          if (CTA2) {
            // one barrier per stage, in the leader CTA: it expects the bytes of both CTAs
            if (crank == 0) mbar_expect_tx(full_bar + s, 2 * STAGE_BYTES);
            const uint32_t lbar = mapa_u32(smem_u32(full_bar + s), 0);
            if (!A_MN) {
              tma_load_2d_pair(sa, ma, lbar, kb * BK, ma0);
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c) tma_load_2d_pair(sa + c * (64 * BK * 2), ma, lbar, ma0 + c * 64, kb * BK);
            }
            if (!B_MN) {
              tma_load_2d_pair(sb, mb, lbar, kb * BK, nb0);
            } else {
#pragma unroll
              for (int c = 0; c < BNL / 64; ++c) tma_load_2d_pair(sb + c * (64 * BK * 2), mb, lbar, nb0 + c * 64, kb * BK);
            }
          } else {
          mbar_expect_tx(full_bar + s, STAGE_BYTES);
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
    constexpr uint32_t idesc = make_idesc(BN, A_MN, B_MN, TM);
    constexpr uint32_t KSTEP_A = A_MN ? 2048u : 32u, KSTEP_B = B_MN ? 2048u : 32u;   // bytes per UMMA_K step
    const uint32_t tiles_addr = smem_u32(tiles);
    const uint64_t da0 = A_MN ? make_desc(tiles_addr, 64 * BK * 2, 1024) : make_desc(tiles_addr, 16, 1024);
    const uint64_t db0 = B_MN ? make_desc(tiles_addr + A_BYTES, 64 * BK * 2, 1024) : make_desc(tiles_addr + A_BYTES, 16, 1024);
    const bool leader = elect_one();
    int s = 0, ui = 0;
    uint32_t ph = 0;
    // CTA pair: only the leader CTA issues (its MMAs read both CTAs' shared memory and write both CTAs' TMEM)
    for (int u; !(CTA2 && crank != 0) && (u = unit_at(ui)) >= 0; ++ui) {
      int m0, n0, kb0, kb1;
      decode(u, m0, n0, kb0, kb1);
      const int a = ui % ACC;
      const uint32_t aph = (ui / ACC) & 1;
      mbar_wait(tmem_empty + a, aph ^ 1);          // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(a * BN);
      uint32_t acc_flag = 0u;
      if (leader) stamp(ui, 2);                    // accumulator stage available
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar + s, ph);
        tc_fence_after();
        if (leader) {
          if (kb == kb0) stamp(ui, 3);             // first operands have landed
          const uint64_t da = da0 + (uint64_t)((uint32_t)s * (STAGE_BYTES >> 4));
          const uint64_t db = db0 + (uint64_t)((uint32_t)s * (STAGE_BYTES >> 4));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            if (CTA2) umma_bf16_pair(tmem_d, da + (uint64_t)(k * (KSTEP_A >> 4)), db + (uint64_t)(k * (KSTEP_B >> 4)), idesc, k == 0 ? acc_flag : 1u);
            else umma_bf16(tmem_d, da + (uint64_t)(k * (KSTEP_A >> 4)), db + (uint64_t)(k * (KSTEP_B >> 4)), idesc, k == 0 ? acc_flag : 1u);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
          if (CTA2) umma_commit_pair(empty_bar + s); else umma_commit(empty_bar + s);
        }
        acc_flag = 1u;
        if (++s == STAGES) { s = 0; ph ^= 1u; }
      }
      if (leader) { if (CTA2) umma_commit_pair(tmem_full + a); else umma_commit(tmem_full + a); }  // accumulator of this unit complete
      if (leader) stamp(ui, 4);                    // last MMA of the unit issued
      __syncwarp();
    }
    if (CTA2 && crank == 0 && ui > 0) {
      // the peer's epilogue warps arrive on this CTA's barriers: see the last hand-back before the barriers go away
      mbar_wait(tmem_empty + ((ui - 1) % ACC), ((ui - 1) / ACC) & 1);
    }
  } else {
    // ===== epilogue: TMEM -> registers -> fused epilogue -> global, one accumulator row per thread =====
    // tcgen05.ld hands each thread 32 consecutive columns of ONE row: 64 bytes of a bf16 output (128 of an fp32 one), i.e.
    // whole 32-byte sectors.  The thread finishes its row piece in registers and stores it with 256-bit (or 128-bit)
    // accesses: no shared-memory transpose, no warp synchronisation, and every term of the fused epilogue that lives in
    // global memory (old C, relu gate, addend) is requested before the TMEM wait.
    constexpr bool E_GEN = EPI == 0, E_RMW = EPI == 0 || EPI == 2, E_F32 = EPI == 0 || EPI == 3, E_BF16 = EPI != 3;
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const size_t esz = E_GEN ? (p.c_f32 ? 4 : 2) : (E_F32 ? 4 : 2);
    // vector path: 16-byte aligned rows for every tensor touched with vector accesses (al32: 32-byte, 256-bit accesses)
    const uintptr_t all_ptrs = reinterpret_cast<uintptr_t>(p.C) | (PAIR ? reinterpret_cast<uintptr_t>(p.s2.C) : 0) | ((size_t)p.ldc * esz) |
                               (p.e.gate ? (reinterpret_cast<uintptr_t>(p.e.gate) | ((size_t)p.e.gate_ld * esz)) : 0) |
                               (p.e.c2 ? (reinterpret_cast<uintptr_t>(p.e.c2) | ((size_t)p.e.c2_ld * esz)) : 0) |
                               ((size_t)p.c_block_off_or * 4);
    const bool vec_ok = !E_GEN || (all_ptrs & 15) == 0 && (!p.e.bias || (reinterpret_cast<uintptr_t>(p.e.bias) & 15) == 0) &&
                        (!p.e.addend || ((reinterpret_cast<uintptr_t>(p.e.addend) | ((size_t)p.e.addend_ld * 4)) & 15) == 0) &&
                        (p.e.alpha_cols % 32 == 0) && (p.c_block_cols % 32 == 0);
    const bool al32 = (all_ptrs & 31) == 0;
    const float alpha0 = (E_GEN && p.e.alpha && p.e.alpha_cols == 0) ? __ldg(p.e.alpha) : 1.f;
    for (int ui = 0, u; (u = unit_at(ui)) >= 0; ++ui) {
      int m0, n0, kb0, kb1;
      decode(u, m0, n0, kb0, kb1);
      m0 += (int)crank * BM;                  // CTA pair: this CTA's TMEM holds rows 128 r .. of the unit
      // the few things in which the launch's second problem differs from the first
      const bool sec = PAIR && u >= units1;
      const int uM = sec ? p.s2.M : p.M;
      void* const uC = sec ? p.s2.C : p.C;
      const int uAcc = E_RMW ? (sec ? p.s2.accumulate : p.e.accumulate) : 0;
      const int acc_stage = ui % ACC;
      const int rbase = m0 + q * 32;
      const int r = rbase + lane;               // this thread's row
      const bool valid = r < uM && !(p.dbg & 1);
      const int rc = min(r, uM - 1);            // clamped (loads only)
      // 32-column blocks this warp really has to move (none if its rows or the unit's K range are empty)
      const int nblk = (rbase < uM && kb1 > kb0) ? min(BN / 32, (p.N - n0 + 31) / 32) : 0;
      constexpr int CSTEP = EPIW / 4;                      // warps sharing a lane quarter interleave the column blocks
      const int cfirst = (warp - 2) >> 2;
      // read-modify-write epilogues (bf16 C): pull this thread's pieces of the old C / gate rows into L2 while the MMAs of this
      // unit are still running
      if (E_RMW && !p.c_f32 && (uAcc || p.e.gate) && valid) {
        for (int cc = cfirst; cc < nblk; cc += CSTEP) {
          const int c = n0 + cc * 32;
          if (uAcc) asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const bf16*>(uC) + (size_t)r * p.ldc + c));
          if (p.e.gate) asm volatile("prefetch.global.L2 [%0];" ::"l"(static_cast<const bf16*>(p.e.gate) + (size_t)r * p.e.gate_ld + c));
        }
      }
      // per-row terms
      const float rs = (E_RMW && p.e.addend && p.e.row_scale) ? __ldg(p.e.row_scale + rc) : 1.f;
      const float* arow = (E_RMW && p.e.addend) ? p.e.addend + (size_t)(rc / p.e.addend_rows) * p.e.addend_ld : nullptr;
      long long r2 = -1;
      if (E_RMW && p.e.c2) {
        const int rr = r % p.e.c2_rows_in;
        if (rr < p.e.c2_rows_keep) r2 = (long long)(r / p.e.c2_rows_in) * p.e.c2_rows_keep + rr;
      }
      // this warp's bias columns of the unit (<= 4 blocks of 32) go to its private shared-memory strip while the MMAs still run:
      // the per-block reads are then broadcast LDS instead of one L2 round trip per block
      float* const bsm = bias_stage + (warp - 2) * 128;
      if (p.e.bias && vec_ok) {
        const int bc = n0 + (cfirst + (lane >> 3) * CSTEP) * 32 + (lane & 7) * 4;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bc + 4 <= p.N && (lane >> 3) * CSTEP + cfirst < BN / 32) bv = __ldg(reinterpret_cast<const float4*>(p.e.bias + bc));
        __syncwarp();                            // the previous unit's reads of the strip are done
        reinterpret_cast<float4*>(bsm)[lane] = bv;
        __syncwarp();
      }
      if (warp == 2 && lane == 0) stamp(ui, 5);    // epilogue ready for this unit
      mbar_wait(tmem_full + acc_stage, (ui / ACC) & 1);
      tc_fence_after();
      if (warp == 2 && lane == 0) stamp(ui, 6);    // accumulator complete
      const uint32_t tmem_acc = tmem_base + (uint32_t)(acc_stage * BN) + ((uint32_t)(q * 32) << 16);
      // Row pointers of this thread for the unit (column n0); the block loop only adds its column offset.  The loop is fully
      // unrolled over the warp's <= 4 blocks with two alternating accumulator register sets, so the next block's tcgen05.ld
      // lands while the current block is processed and nothing is copied.  The epilogue is bound by instruction issue (two
      // warps per scheduler, dependent chains), hence packed fp32x2 arithmetic, relu folded into the bf16 conversion and
      // the relu gate applied to the packed result.
      unsigned char* const crow = static_cast<unsigned char*>(uC) + ((size_t)rc * p.ldc + n0) * esz;
      unsigned char* const c2row = r2 >= 0 ? static_cast<unsigned char*>(p.e.c2) + ((size_t)r2 * p.e.c2_ld + n0) * esz : nullptr;
      const unsigned char* const grow = (E_RMW && p.e.gate) ? static_cast<const unsigned char*>(p.e.gate) + ((size_t)rc * p.e.gate_ld + n0) * esz : nullptr;
      const float* const arow0 = arow ? arow + n0 : nullptr;
      const bool f_relu = (E_GEN || EPI == 1) && p.e.relu != 0, f_bias = p.e.bias != nullptr, f_alpha = E_GEN && p.e.alpha != nullptr;
      const bool f_f32 = E_GEN ? p.c_f32 != 0 : EPI == 3;
      constexpr int NB = (BN / 32) / CSTEP;                // blocks per warp and unit
      uint32_t accA[32], accB[32];
      auto block = [&](uint32_t (&x)[32], uint32_t (&nxt)[32], const int cc, const int slot) {
        const int c = n0 + cc * 32;                        // first of this thread's 32 columns
        const int co = cc * 32;                            // ... relative to the row pointers
        const bool vec = vec_ok && c + 32 <= p.N;
        // terms from global memory: requested before the TMEM wait so that their latency is hidden
        // (one 128-byte register buffer: the old C + gate pieces of a bf16 read-modify-write epilogue, or the addend piece)
        const float al = (vec && f_alpha && p.e.alpha_cols) ? __ldg(p.e.alpha + c / p.e.alpha_cols) : alpha0;
        const bool pre_acc = E_RMW && vec && !f_f32 && uAcc, pre_gate = E_RMW && vec && !f_f32 && grow;
        const bool pre_add = E_RMW && vec && arow0 && !pre_acc && !pre_gate, late_add = E_RMW && vec && arow0 && !pre_add;
        uint4 pre[8];
        if (pre_acc) {
          const uint4* src = reinterpret_cast<const uint4*>(crow + (size_t)co * 2);
#pragma unroll
          for (int g = 0; g < 4; ++g) pre[g] = ldg_v4u(src + g);
        }
        if (pre_gate) {
          const uint4* src = reinterpret_cast<const uint4*>(grow + (size_t)co * 2);
#pragma unroll
          for (int g = 0; g < 4; ++g) pre[4 + g] = ldg_nc_v4u(src + g);
        }
        if (pre_add) {
#pragma unroll
          for (int g = 0; g < 8; ++g) pre[g] = ldg_nc_v4u(arow0 + co + 4 * g);
        }
        tmem_ld32_wait(x);
        if (cc + CSTEP < nblk) {
          if (!(p.dbg & 2)) tmem_ld32_issue(tmem_acc + (uint32_t)((cc + CSTEP) * 32), nxt);   // overlaps the work below
        } else {
          // every TMEM read of this unit has completed: hand the accumulator stage back to the MMA issuer early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) { if (CTA2) mbar_arrive_cluster(mapa_u32(smem_u32(tmem_empty + acc_stage), 0)); else mbar_arrive(tmem_empty + acc_stage); }
        }
        if (E_GEN && !vec) {
          // ragged edge / unaligned tensors: generic per-element path (kept out of line: it must not bloat the hot loop)
          if (valid) {
#pragma unroll
            for (int g = 0; g < 8; ++g)      // unrolled: x[] must stay in registers
              if (c + 4 * g < p.N)
                epi_slow(p, uC, uAcc, r, c + 4 * g, make_float4(__uint_as_float(x[4 * g]), __uint_as_float(x[4 * g + 1]),
                                                              __uint_as_float(x[4 * g + 2]), __uint_as_float(x[4 * g + 3])));
          }
          return;
        }
        if (pre_add) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            x[4 * g] = __float_as_uint(fmaf(rs, __uint_as_float(pre[g].x), __uint_as_float(x[4 * g])));
            x[4 * g + 1] = __float_as_uint(fmaf(rs, __uint_as_float(pre[g].y), __uint_as_float(x[4 * g + 1])));
            x[4 * g + 2] = __float_as_uint(fmaf(rs, __uint_as_float(pre[g].z), __uint_as_float(x[4 * g + 2])));
            x[4 * g + 3] = __float_as_uint(fmaf(rs, __uint_as_float(pre[g].w), __uint_as_float(x[4 * g + 3])));
          }
        } else if (late_add) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 a = ldg_nc_v4(arow0 + co + 4 * g);
            x[4 * g] = __float_as_uint(fmaf(rs, a.x, __uint_as_float(x[4 * g])));
            x[4 * g + 1] = __float_as_uint(fmaf(rs, a.y, __uint_as_float(x[4 * g + 1])));
            x[4 * g + 2] = __float_as_uint(fmaf(rs, a.z, __uint_as_float(x[4 * g + 2])));
            x[4 * g + 3] = __float_as_uint(fmaf(rs, a.w, __uint_as_float(x[4 * g + 3])));
          }
        }
        if (f_bias) {
          const float4* bsrc = reinterpret_cast<const float4*>(bsm + slot * 32);
          if (f_alpha) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 bv = bsrc[g];                                                    // warp-uniform address: broadcast
              ffma2(x[4 * g], x[4 * g + 1], al, bv.x, bv.y); ffma2(x[4 * g + 2], x[4 * g + 3], al, bv.z, bv.w);
            }
          } else {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              const float4 bv = bsrc[g];
              fadd2(x[4 * g], x[4 * g + 1], bv.x, bv.y); fadd2(x[4 * g + 2], x[4 * g + 3], bv.z, bv.w);
            }
          }
        } else if (f_alpha) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) ffma2(x[j], x[j + 1], al, 0.f, 0.f);
        }
        if (E_F32 && f_f32) {
          if (f_relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = __float_as_uint(fmaxf(__uint_as_float(x[j]), 0.f));
          }
          // column-block scatter (weight gradients of side-by-side layers): block j of C lives at C + c_block_off[j]
          float* dst = reinterpret_cast<float*>(crow) + co;
          if (p.c_block_cols) { const int jb = c / p.c_block_cols; dst += p.c_block_off[jb] - (long long)jb * p.c_block_cols; }
          if (p.atomic_out) {
            if (valid) {
#pragma unroll
              for (int g = 0; g < 8; ++g)
                red_add_v4(dst + 4 * g, make_float4(__uint_as_float(x[4 * g]), __uint_as_float(x[4 * g + 1]), __uint_as_float(x[4 * g + 2]),
                                                    __uint_as_float(x[4 * g + 3])));
            }
          } else {
            if (E_GEN && uAcc) {
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 o = ldg_v4(dst + 4 * g);
                fadd2(x[4 * g], x[4 * g + 1], o.x, o.y); fadd2(x[4 * g + 2], x[4 * g + 3], o.z, o.w);
              }
            }
            if (E_GEN && grow) {
              const float* g0 = reinterpret_cast<const float*>(grow) + co;
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 gv = ldg_nc_v4(g0 + 4 * g);
                x[4 * g] = gv.x > 0.f ? x[4 * g] : 0u; x[4 * g + 1] = gv.y > 0.f ? x[4 * g + 1] : 0u;
                x[4 * g + 2] = gv.z > 0.f ? x[4 * g + 2] : 0u; x[4 * g + 3] = gv.w > 0.f ? x[4 * g + 3] : 0u;
              }
            }
            if (valid) {
              if (al32) {
#pragma unroll
                for (int g = 0; g < 4; ++g) stg_v8u(dst + 8 * g, x + 8 * g);
              } else {
#pragma unroll
                for (int g = 0; g < 8; ++g) stg_v4u(dst + 4 * g, x + 4 * g);
              }
              if (E_GEN && c2row) {
                float* d2 = reinterpret_cast<float*>(c2row) + co;
#pragma unroll
                for (int g = 0; g < 8; ++g) stg_v4u(d2 + 4 * g, x + 4 * g);
              }
            }
          }
        } else if (E_BF16) {
          uint32_t w[16];
          if (f_relu && !pre_acc) {
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j] = pack_bf16_relu(__uint_as_float(x[2 * j]), __uint_as_float(x[2 * j + 1]));
          } else {
            if (f_relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) x[j] = __float_as_uint(fmaxf(__uint_as_float(x[j]), 0.f));
            }
            if (pre_acc) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                fadd2(x[8 * g], x[8 * g + 1], bf_lo(pre[g].x), bf_hi(pre[g].x)); fadd2(x[8 * g + 2], x[8 * g + 3], bf_lo(pre[g].y), bf_hi(pre[g].y));
                fadd2(x[8 * g + 4], x[8 * g + 5], bf_lo(pre[g].z), bf_hi(pre[g].z)); fadd2(x[8 * g + 6], x[8 * g + 7], bf_lo(pre[g].w), bf_hi(pre[g].w));
              }
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j] = pack_bf16(__uint_as_float(x[2 * j]), __uint_as_float(x[2 * j + 1]));
          }
          if (pre_gate) {
            // gate > 0 per bf16 half, applied to the packed result (a rejected element becomes +0)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              w[4 * g] &= bf16x2_gt0_mask(pre[4 + g].x); w[4 * g + 1] &= bf16x2_gt0_mask(pre[4 + g].y);
              w[4 * g + 2] &= bf16x2_gt0_mask(pre[4 + g].z); w[4 * g + 3] &= bf16x2_gt0_mask(pre[4 + g].w);
            }
          }
          if (valid) {
            bf16* dst = reinterpret_cast<bf16*>(crow) + co;
            if (al32) { stg_v8u(dst, w); stg_v8u(dst + 16, w + 8); }
            else {
#pragma unroll
              for (int g = 0; g < 4; ++g) stg_v4u(dst + 8 * g, w + 4 * g);
            }
            if (E_RMW && c2row) {
              bf16* d2 = reinterpret_cast<bf16*>(c2row) + co;
              if (al32) { stg_v8u(d2, w); stg_v8u(d2 + 16, w + 8); }
              else {
#pragma unroll
                for (int g = 0; g < 4; ++g) stg_v4u(d2 + 8 * g, w + 4 * g);
              }
            }
          }
        }
      };
      if (cfirst < nblk && !(p.dbg & 2)) tmem_ld32_issue(tmem_acc + (uint32_t)(cfirst * 32), accA);
      if constexpr (EPI == 1 || EPI == 3) {
        // lean kinds: all blocks unrolled
#pragma unroll
        for (int i = 0; i < NB; i += 2) {
          const int cc = cfirst + i * CSTEP;
          if (cc < nblk) block(accA, accB, cc, i);
          if (cc + CSTEP < nblk) block(accB, accA, cc + CSTEP, i + 1);
        }
      } else {
        // unrolled by two -- one round trip through the two register sets (all four would need > 168 registers, the most a
        // 10-warp CTA can have)
#pragma unroll 1
        for (int i = 0; i < NB; i += 2) {
          const int cc = cfirst + i * CSTEP;
          if (cc < nblk) block(accA, accB, cc, i);
          if (cc + CSTEP < nblk) block(accB, accA, cc + CSTEP, i + 1);
        }
      }
      if (cfirst >= nblk) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if (CTA2) mbar_arrive_cluster(mapa_u32(smem_u32(tmem_empty + acc_stage), 0)); else mbar_arrive(tmem_empty + acc_stage); }
      }
      if (warp == 2 && lane == 0) stamp(ui, 7);    // unit stored
    }  // units
  }
  if (tr && threadIdx.x == 0) tr[1] = clock64();
  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();     // pair: neither CTA's shared memory / TMEM goes away under the other
  if (warp == 1) {
    tc_fence_after();
    if (CTA2) tmem_dealloc_pair(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (tr && threadIdx.x == 0) tr[2] = clock64();
}

// ================================================================== dependent chain of batch-sized products in ONE launch
// The classifier / BUTD tail of the step (fusion.py:37-52, classifier.py:14-25, train.py:107-108 and their transposes) is a
// chain of seven products whose M is the batch (256 rows): each depends on the previous one, each is < 1 % of the step's
// FLOPs, and launched one by one each costs a launch, a pipeline fill and a drain (8-22 us a piece, ~120 us together, with
// 24-98 CTAs on 148 SMs).  Here they are stages of one persistent launch: every CTA keeps its barriers, TMEM and warp roles,
// runs its tiles of stage i, and a grid-wide arrival counter (release / acquire at gpu scope, then fence.proxy.async before
// the next stage's TMA reads what other SMs' epilogues stored) separates the stages.  128 x 64 tiles, 8-deep ring, operand
// major-ness per stage at run time (forward: weights MN-major, input gradients: K-major).  The loss (BCE + score + dlogits)
// is a stage of its own executed by the epilogue warps.  All CTAs of the launch must be resident together: grid <= SM count.
constexpr int CH_BN = 64, CH_STAGES = 8, CH_ACC = 2, CH_EPIW = 4, CH_MAXP = 8;
struct ChainProblem {
  int kind;                       // 0 product, 1 loss
  int M, N, total_kb, tiles_n, units, a_mn, b_mn;
  void* C; int ldc, c_f32;        // C = f(acc) [* mul]
  const float* bias; int relu;
  const bf16* gate; int gate_ld;  // keep where gate > 0
  const bf16* mul; int mul_ld;    // C = x * mul
  bf16* out2; int out2_ld; const bf16* mul2; int mul2_ld;   // out2 = x * mul2 (bf16)
  // loss stage (train.py:20-26,107-108): logits [M, ldl] fp32, target [M, A] fp32 -> loss, score, dlog [M, ldd] bf16
  const float* logits; int ldl; const float* target; int A; float inv_B, gscale; float* loss; float* score; bf16* dlog; int ldd;
};
struct ChainParams {
  CUtensorMap map[2 * CH_MAXP];
  ChainProblem prob[CH_MAXP];
  int nprob;
  unsigned int* counter;          // grid arrival counter (0 between launches)
  long long* trace;               // measurement aid (regat_gemm_trace): CTA 0's clock64 at kernel start [128], when a stage's wait is passed [152 + stage] and at its end [136 + stage]
};

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// warp-collective: lane 0 polls (one L2 transaction per warp and round), everybody leaves together; the trailing fence orders the
// other lanes' later reads behind lane 0's acquire
__device__ __forceinline__ void chain_wait(const unsigned int* ctr, unsigned int need) {
  if ((threadIdx.x & 31) == 0) {
    uint32_t spin = 0;
#pragma unroll 1
    while (ld_acquire_gpu(ctr) < need) {
      if (++spin > SPIN_LIMIT) __trap();
      __nanosleep(32);
    }
  }
  __syncwarp();
  __threadfence();
}
__device__ __forceinline__ uint4 ldg_cg_v4u(const void* p) {     // L2 only: the data was written by other SMs earlier in this launch
  uint4 v;
  asm volatile("ld.global.cg.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(64 + 32 * CH_EPIW, 1) gemm_chain_kernel(const __grid_constant__ ChainParams cp) {
  constexpr int BN = CH_BN, STAGES = CH_STAGES, ACC = CH_ACC, EPIW = CH_EPIW;
  constexpr uint32_t A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = ACC * BN;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* tiles = smem;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + ACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + ACC);
  float* red = reinterpret_cast<float*>(tmem_slot + 4);       // 16 floats: reductions of the loss stage
  float* bias_stage = red + 16;                                // EPIW x 64 floats: each epilogue warp's bias columns of a unit
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x;
  if (threadIdx.x < 2 * cp.nprob && cp.prob[threadIdx.x >> 1].kind == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&cp.map[threadIdx.x]) : "memory");
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
    for (int a = 0; a < ACC; ++a) { mbar_init(tmem_full + a, 1); mbar_init(tmem_empty + a, EPIW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long* const tr = (cp.trace && blockIdx.x == 0) ? cp.trace : nullptr;
  if (tr && threadIdx.x == 0) tr[128] = clock64();

  if (warp == 0) {
    // ===== TMA producer
    const bool leader = elect_one();
    int s = 0;
    uint32_t ph = 0;
    for (int pi = 0; pi < cp.nprob; ++pi) {
      const ChainProblem& P = cp.prob[pi];
      if (P.kind != 0) continue;
      if (pi > 0) {
        // every CTA's epilogue of the stages before has stored (and released) its results; what follows reads them through TMA
        chain_wait(cp.counter, (unsigned)G * (unsigned)pi);
        asm volatile("fence.proxy.async;" ::: "memory");
      }
      const CUtensorMap* ma = &cp.map[2 * pi];
      const CUtensorMap* mb = &cp.map[2 * pi + 1];
      for (int u = blockIdx.x; u < P.units; u += G) {
        const int tm = u / P.tiles_n, tn = u - tm * P.tiles_n;
        const int m0 = tm * BM, n0 = tn * BN;
        for (int kb = 0; kb < P.total_kb; ++kb) {
          mbar_wait(empty_bar + s, ph ^ 1);
          if (leader) {
            mbar_expect_tx(full_bar + s, STAGE_BYTES);
            unsigned char* sa = tiles + s * STAGE_BYTES;
            unsigned char* sb = sa + A_BYTES;
            if (!P.a_mn) {
              tma_load_2d(sa, ma, full_bar + s, kb * BK, m0);
            } else {
#pragma unroll
              for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * (64 * BK * 2), ma, full_bar + s, m0 + c * 64, kb * BK);
            }
            if (!P.b_mn) tma_load_2d(sb, mb, full_bar + s, kb * BK, n0);
            else tma_load_2d(sb, mb, full_bar + s, n0, kb * BK);
          }
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    const uint32_t tiles_addr = smem_u32(tiles);
    const bool leader = elect_one();
    int s = 0, ui = 0;
    uint32_t ph = 0;
    for (int pi = 0; pi < cp.nprob; ++pi) {
      const ChainProblem& P = cp.prob[pi];
      if (P.kind != 0) continue;
      const uint32_t idesc = make_idesc(BN, P.a_mn != 0, P.b_mn != 0, BM);
      const uint32_t kstep_a = P.a_mn ? (2048u >> 4) : (32u >> 4), kstep_b = P.b_mn ? (2048u >> 4) : (32u >> 4);
      const uint64_t da0 = P.a_mn ? make_desc(tiles_addr, 64 * BK * 2, 1024) : make_desc(tiles_addr, 16, 1024);
      const uint64_t db0 = P.b_mn ? make_desc(tiles_addr + A_BYTES, 64 * BK * 2, 1024) : make_desc(tiles_addr + A_BYTES, 16, 1024);
      for (int u = blockIdx.x; u < P.units; u += G, ++ui) {
        const int a = ui % ACC;
        mbar_wait(tmem_empty + a, ((ui / ACC) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(a * BN);
        uint32_t acc_flag = 0u;
        for (int kb = 0; kb < P.total_kb; ++kb) {
          mbar_wait(full_bar + s, ph);
          tc_fence_after();
          if (leader) {
            const uint64_t da = da0 + (uint64_t)((uint32_t)s * (STAGE_BYTES >> 4));
            const uint64_t db = db0 + (uint64_t)((uint32_t)s * (STAGE_BYTES >> 4));
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              umma_bf16(tmem_d, da + (uint64_t)(k * kstep_a), db + (uint64_t)(k * kstep_b), idesc, k == 0 ? acc_flag : 1u);
            umma_commit(empty_bar + s);
          }
          acc_flag = 1u;
          if (++s == STAGES) { s = 0; ph ^= 1u; }
        }
        if (leader) umma_commit(tmem_full + a);
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue warps: products' epilogues, the loss stage, and the grid arrival of every stage
    const int q = warp & 3;
    const int et = (warp - 2) * 32 + lane;           // 0..127
    int ui = 0;
    for (int pi = 0; pi < cp.nprob; ++pi) {
      const ChainProblem& P = cp.prob[pi];
      if (pi > 0) chain_wait(cp.counter, (unsigned)G * (unsigned)pi);     // operands of this stage's epilogue / loss are visible
      if (tr && et == 0) tr[152 + pi] = clock64();
      if (P.kind == 0) {
        for (int u = blockIdx.x; u < P.units; u += G, ++ui) {
          const int tm = u / P.tiles_n, tn = u - tm * P.tiles_n;
          const int m0 = tm * BM, n0 = tn * BN;
          const int a = ui % ACC;
          const int r = m0 + q * 32 + lane;
          const bool valid = r < P.M;
          const int rc = min(r, P.M - 1);
          // Everything the epilogue reads from global memory is requested BEFORE the accumulator wait (bias: this warp's 64
          // columns into its shared-memory strip; gate / mul / mul2: this thread's 2 x 64-byte row pieces into registers), so
          // the only latency left after the last MMA is the TMEM read.
          float* const bsm = bias_stage + (warp - 2) * 64;
          if (P.bias) {
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane < 16) {
              const int bc = n0 + 4 * lane;
              if (bc + 4 <= P.N) bv = __ldg(reinterpret_cast<const float4*>(P.bias + bc));
              else {
                if (bc < P.N) bv.x = __ldg(P.bias + bc);
                if (bc + 1 < P.N) bv.y = __ldg(P.bias + bc + 1);
                if (bc + 2 < P.N) bv.z = __ldg(P.bias + bc + 2);
              }
            }
            __syncwarp();
            if (lane < 16) reinterpret_cast<float4*>(bsm)[lane] = bv;
            __syncwarp();
          }
          uint4 pg[2][4], pm[2][4], pm2[2][4];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int c = n0 + half * 32;
            if (c + 32 > P.N) continue;       // gate / mul terms exist for full blocks only (host-checked: N % 32 == 0)
            if (P.gate) {
              const uint4* g4 = reinterpret_cast<const uint4*>(P.gate + (size_t)rc * P.gate_ld + c);
#pragma unroll
              for (int g = 0; g < 4; ++g) pg[half][g] = ldg_cg_v4u(g4 + g);
            }
            if (P.mul) {
              const uint4* m4 = reinterpret_cast<const uint4*>(P.mul + (size_t)rc * P.mul_ld + c);
#pragma unroll
              for (int g = 0; g < 4; ++g) pm[half][g] = ldg_cg_v4u(m4 + g);
            }
            if (P.out2) {
              const uint4* m4 = reinterpret_cast<const uint4*>(P.mul2 + (size_t)rc * P.mul2_ld + c);
#pragma unroll
              for (int g = 0; g < 4; ++g) pm2[half][g] = ldg_cg_v4u(m4 + g);
            }
          }
          mbar_wait(tmem_full + a, (ui / ACC) & 1);
          tc_fence_after();
          const uint32_t tmem_acc = tmem_base + (uint32_t)(a * BN) + ((uint32_t)(q * 32) << 16);
          uint32_t acc0[32], acc1[32];
          tmem_ld32_issue(tmem_acc, acc0);
          tmem_ld32_issue(tmem_acc + 32u, acc1);
          tmem_ld32_wait(acc0);
          tmem_ld32_wait(acc1);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tmem_empty + a);       // the accumulator stage is free again
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t (&x)[32] = half ? acc1 : acc0;
            const int c = n0 + half * 32;
            if (c >= P.N) continue;
            const bool full = c + 32 <= P.N;
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(x[j]);
            if (P.bias) {
              const float4* bsrc = reinterpret_cast<const float4*>(bsm + half * 32);
#pragma unroll
              for (int g = 0; g < 8; ++g) {
                const float4 bv = bsrc[g];
                v[4 * g] += bv.x; v[4 * g + 1] += bv.y; v[4 * g + 2] += bv.z; v[4 * g + 3] += bv.w;
              }
            }
            if (P.relu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (P.gate && full) {
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint32_t ws[4] = {pg[half][g].x, pg[half][g].y, pg[half][g].z, pg[half][g].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if (!(bf_lo(ws[k]) > 0.f)) v[8 * g + 2 * k] = 0.f;
                  if (!(bf_hi(ws[k]) > 0.f)) v[8 * g + 2 * k + 1] = 0.f;
                }
              }
            }
            if (P.out2 && full) {      // out2 = x * mul2
              uint32_t w2[16];
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint32_t ws[4] = {pm2[half][g].x, pm2[half][g].y, pm2[half][g].z, pm2[half][g].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) w2[4 * g + k] = pack_bf16(v[8 * g + 2 * k] * bf_lo(ws[k]), v[8 * g + 2 * k + 1] * bf_hi(ws[k]));
              }
              if (valid) {
                bf16* d2 = P.out2 + (size_t)r * P.out2_ld + c;
#pragma unroll
                for (int g = 0; g < 4; ++g) stg_v4u(d2 + 8 * g, w2 + 4 * g);
              }
            }
            if (P.mul && full) {       // C = x * mul
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const uint32_t ws[4] = {pm[half][g].x, pm[half][g].y, pm[half][g].z, pm[half][g].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) { v[8 * g + 2 * k] *= bf_lo(ws[k]); v[8 * g + 2 * k + 1] *= bf_hi(ws[k]); }
              }
            }
            if (valid) {
              if (P.c_f32) {
                float* dst = static_cast<float*>(P.C) + (size_t)r * P.ldc + c;
                if (full) {
#pragma unroll
                  for (int g = 0; g < 8; ++g) stg_v4u(dst + 4 * g, reinterpret_cast<const uint32_t*>(v) + 4 * g);
                } else {
#pragma unroll
                  for (int j = 0; j < 32; ++j)
                    if (c + j < P.N) dst[j] = v[j];
                }
              } else {
                bf16* dst = static_cast<bf16*>(P.C) + (size_t)r * P.ldc + c;
                uint32_t w[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) w[j] = pack_bf16(v[2 * j], v[2 * j + 1]);
#pragma unroll
                for (int g = 0; g < 4; ++g) stg_v4u(dst + 8 * g, w + 4 * g);
              }
            }
          }
        }
      } else {
        // ----- loss stage: one row per CTA and round, 128 threads per row
        for (int b = blockIdx.x; b < P.M; b += G) {
          const float* xrow = P.logits + (size_t)b * P.ldl;
          const float* zrow = P.target + (size_t)b * P.A;
          float ls = 0.f, bv = -INFINITY;
          int bi = 0x7fffffff;
          // batches of 8 elements per thread: all sixteen loads of a batch are in flight before the first use (with 4 warps per
          // row there is no other latency hiding)
          constexpr int LB = 8;
          for (int a0 = et; a0 < P.ldd; a0 += LB * 32 * EPIW) {
            float xs[LB], zs[LB];
#pragma unroll
            for (int i = 0; i < LB; ++i) {
              const int a = a0 + i * 32 * EPIW;
              const bool in = a < P.A;
              xs[i] = in ? __ldcg(xrow + a) : 0.f;
              zs[i] = in ? __ldg(zrow + a) : 0.f;
            }
#pragma unroll
            for (int i = 0; i < LB; ++i) {
              const int a = a0 + i * 32 * EPIW;
              if (a >= P.ldd) break;
              float g = 0.f;
              if (a < P.A) {
                const float xv = xs[i], zv = zs[i];
                // tf.nn.sigmoid_cross_entropy_with_logits: max(x, 0) - x z + log1p(exp(-|x|)), one SFU exponential per element:
                // t = exp(-|x|); log1p(t) by its series where 1 + t would round t away; sigmoid(x) = 1/(1+t) or t/(1+t)
                const float t = __expf(-fabsf(xv));
                const float l1p = t < 1e-3f ? t * (1.f - t * (0.5f - t * (1.f / 3.f))) : __logf(1.f + t);
                ls += fmaxf(xv, 0.f) - xv * zv + l1p;
                const float sg = __fdividef(xv >= 0.f ? 1.f : t, 1.f + t);
                g = (sg - zv) * P.inv_B * P.gscale;
                if (xv > bv) { bv = xv; bi = a; }
              }
              P.dlog[(size_t)b * P.ldd + a] = __float2bfloat16_rn(g);
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {        // argmax with first-index tie-break (np.argmax), and the loss sum
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            ls += __shfl_xor_sync(0xffffffffu, ls, o);
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");      // previous row's reads of `red` are done
          if (lane == 0) { red[warp - 2] = ls; red[4 + warp - 2] = bv; reinterpret_cast<int*>(red)[8 + warp - 2] = bi; }
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (et == 0) {
            float tot = 0.f, best = red[4];
            int besti = reinterpret_cast<int*>(red)[8];
            for (int w = 0; w < EPIW; ++w) {
              tot += red[w];
              const float ov = red[4 + w];
              const int oi = reinterpret_cast<int*>(red)[8 + w];
              if (ov > best || (ov == best && oi < besti)) { best = ov; besti = oi; }
            }
            atomicAdd(P.loss, tot * P.inv_B);
            if (P.score) atomicAdd(P.score, __ldg(zrow + besti));
          }
        }
      }
      // ----- stage done on this CTA: publish
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
        if (tr) tr[136 + pi] = clock64();
        const unsigned int old = atomicAdd(cp.counter, 1u);
        if (pi == cp.nprob - 1 && old == (unsigned)G * (unsigned)cp.nprob - 1u) {
          // last arrival of the launch: every CTA has passed every wait -- leave the counter at 0 for the next launch
          atomicExch(cp.counter, 0u);
        }
      }
    }
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

template <int BN, int STAGES, int ACC, int EPIW, bool CTA2, bool AM, bool BMN, bool PR, int EPI>
int launch_one(const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, const CUtensorMap& ma2, const CUtensorMap& mb2, int ctas,
               cudaStream_t st) {
  constexpr int BNL = CTA2 ? BN / 2 : BN;
  constexpr size_t smem = (size_t)STAGES * (BM * BK * 2 + BNL * BK * 2) + (2 * STAGES + 2 * ACC) * 8 + 16 + EPIW * 512;
  auto kern = gemm_tc_kernel<BN, STAGES, ACC, EPIW, AM, BMN, PR, CTA2, EPI>;
  // the shared-memory attribute is per (function, device): one flag per device, set under a lock
  static std::mutex mu;
  static bool attr_set[64] = {};
  int dev = 0;
  REGAT_CUDA(cudaGetDevice(&dev));
  {
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      REGAT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
  }
  {
    static const int pdl = [] { const char* s = getenv("REGAT_TC_PDL"); return s ? atoi(s) : 1; }();
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)ctas, 1, 1); cfg.blockDim = dim3(64 + 32 * EPIW, 1, 1);
    cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (CTA2) {
      at[na].id = cudaLaunchAttributeClusterDimension;
      at[na].val.clusterDim.x = 2; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
      ++na;
    }
    if (pdl && pdl_enabled()) {
      at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    REGAT_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, p, ma2, mb2));
  }
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}

// two-problem launches exist for the 256-wide tiles only, and only with the kinds the engine pairs (2) or the generic one
template <int BN, int STAGES, int ACC, int EPIW, bool CTA2, bool AM, bool BMN>
int launch_majors(bool pair, int epi, const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, const CUtensorMap& ma2,
                  const CUtensorMap& mb2, int ctas, cudaStream_t st) {
  if (pair) {
    if constexpr (BN == 256) {
      if (epi == 2) return launch_one<BN, STAGES, ACC, EPIW, CTA2, AM, BMN, true, 2>(ma, mb, p, ma2, mb2, ctas, st);
      return launch_one<BN, STAGES, ACC, EPIW, CTA2, AM, BMN, true, 0>(ma, mb, p, ma2, mb2, ctas, st);
    } else {
      REGAT_REQUIRE(false, REGAT_ERR_UNSUPPORTED, "gemm_tc: two-problem launches need 256-wide tiles");
    }
  }
  switch (epi) {
    case 1: return launch_one<BN, STAGES, ACC, EPIW, CTA2, AM, BMN, false, 1>(ma, mb, p, ma2, mb2, ctas, st);
    case 2: return launch_one<BN, STAGES, ACC, EPIW, CTA2, AM, BMN, false, 2>(ma, mb, p, ma2, mb2, ctas, st);
    case 3: return launch_one<BN, STAGES, ACC, EPIW, CTA2, AM, BMN, false, 3>(ma, mb, p, ma2, mb2, ctas, st);
    default: return launch_one<BN, STAGES, ACC, EPIW, CTA2, AM, BMN, false, 0>(ma, mb, p, ma2, mb2, ctas, st);
  }
}

template <int BN, int STAGES, int ACC, int EPIW, bool CTA2 = false>
int launch_cfg(bool a_mn, bool b_mn, int epi, const CUtensorMap& ma, const CUtensorMap& mb, const TcParams& p, const CUtensorMap& ma2,
               const CUtensorMap& mb2, int ctas, cudaStream_t st) {
  const bool pair = p.s2.units > 0;
  if (!a_mn && b_mn) return launch_majors<BN, STAGES, ACC, EPIW, CTA2, false, true>(pair, epi, ma, mb, p, ma2, mb2, ctas, st);
  if (!a_mn && !b_mn) return launch_majors<BN, STAGES, ACC, EPIW, CTA2, false, false>(pair, epi, ma, mb, p, ma2, mb2, ctas, st);
  if (a_mn && b_mn) return launch_majors<BN, STAGES, ACC, EPIW, CTA2, true, true>(pair, epi, ma, mb, p, ma2, mb2, ctas, st);
  return launch_majors<BN, STAGES, ACC, EPIW, CTA2, true, false>(pair, epi, ma, mb, p, ma2, mb2, ctas, st);
}

// Smallest epilogue kind that covers a call (see the kernel's EPI parameter); `acc2`: a paired second problem accumulates.
int epi_kind(const TcParams& p, bool pair, int acc2, const void* C2) {
  const EpiArgs& e = p.e;
  static const int force = [] { const char* s = getenv("REGAT_TC_EPI"); return s ? atoi(s) : -1; }();
  if (force == 0) return 0;
  const size_t esz = p.c_f32 ? 4 : 2;
  uintptr_t a = reinterpret_cast<uintptr_t>(p.C) | reinterpret_cast<uintptr_t>(C2) | ((size_t)p.ldc * esz) | ((size_t)p.c_block_off_or * 4);
  if (e.gate) a |= reinterpret_cast<uintptr_t>(e.gate) | ((size_t)e.gate_ld * esz);
  if (e.c2) a |= reinterpret_cast<uintptr_t>(e.c2) | ((size_t)e.c2_ld * esz);
  if (e.bias) a |= reinterpret_cast<uintptr_t>(e.bias);
  if (e.addend) a |= reinterpret_cast<uintptr_t>(e.addend) | ((size_t)e.addend_ld * 4);
  if ((a & 15) != 0 || p.N % 32 != 0 || e.alpha || p.c_block_cols % 32 != 0) return 0;
  const bool rmw = e.addend || e.gate || e.c2 || e.accumulate || acc2;
  if (p.c_f32) return (rmw || e.relu || pair) ? 0 : 3;
  return rmw ? 2 : (pair ? 2 : 1);
}

long long* g_trace = nullptr;   // regat_gemm_trace

struct Prepared {
  CUtensorMap ma, mb;
  TcParams p;
  bool a_mn, b_mn;
  int bn;
  bool cta2;     // 256-row units on CTA pairs (cta_group::2)
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
  int tiles_m = ceil_div(M, BM);
  int bn = (force_bn == 64 || force_bn == 128 || force_bn == 256) ? force_bn : ((N >= 256 && tiles_m * ceil_div(N, 256) >= num_sms()) ? 256 : 128);
  // skinny problems (a batch-sized M, long K): the 128-wide tiling leaves most SMs idle behind a serial K loop that is
  // bound by TMA latency at 3 stages -- narrower tiles double the CTAs and the 24 KB stages allow an 8-deep ring
  static const int skinny = [] { const char* s = getenv("REGAT_TC_SKINNY"); return s ? atoi(s) : 1; }();
  if (!force_bn && skinny && bn == 128 && tiles_m <= 2 && tiles_m * ceil_div(N, 128) * 2 <= num_sms() && K >= 512) bn = 64;
  // 256-wide tiles run on CTA pairs: a unit is 256 x 256, each CTA of the pair loads half of it (REGAT_TC_CTA2=0: single CTAs)
  static const int cta2_env = [] { const char* s = getenv("REGAT_TC_CTA2"); return s ? atoi(s) : 1; }();
  // Long-K products with few output tiles (the weight gradients: K = 5120 .. 9216, 16 .. 64 units of 256 x 256): the 128 x 128 tiling
  // with two CTAs per SM pulls 125 bytes per clock and SM through L2; 256 x 256 units on CTA pairs need half of that, and a split-K
  // factor that fills ONE round of pairs (units x splits <= SMs / 2) keeps the red traffic at 1 .. 4 passes over the output
  // (measured, isolated: 39.9 -> 35.3, 38.6 -> 34.1, 39.3 -> 35.4, 25.6 -> 24.4 us; REGAT_TC_WGRAD256=0 keeps the old choice)
  static const int wgrad256 = [] { const char* s = getenv("REGAT_TC_WGRAD256"); return s ? atoi(s) : 1; }();
  const bool plain_out = c_dtype == REGAT_F32 && !e.bias && !e.addend && !e.relu && !e.gate && !e.c2 && !e.accumulate;
  int pair_splits = 0;
  if (!force_bn && wgrad256 && cta2_env && plain_out && split_k <= 1 && K >= 2048 && M >= 256 && N >= 256) {
    const int t256 = ceil_div(M, 2 * BM) * ceil_div(N, 256), pairs = std::max(1, num_sms() / 2);
    if (t256 <= pairs) { bn = 256; pair_splits = std::max(1, pairs / t256); }
  }
  const bool cta2 = bn == 256 && cta2_env != 0;
  if (cta2) tiles_m = ceil_div(M, 2 * BM);
  const int tiles_n = ceil_div(N, bn);
  const int total_kb = ceil_div(K, BK);
  // split-K only for plain fp32 targets (weight gradients: few output tiles, very long K): fill ~2 CTAs per SM
  int splits = 1;
  const bool plain = c_dtype == REGAT_F32 && !e.bias && !e.addend && !e.relu && !e.gate && !e.c2 && !e.accumulate;
  static const int auto_split = [] { const char* s = getenv("REGAT_TC_SPLITK"); return s ? atoi(s) : 1; }();
  if (plain) {
    static const int force_splits = [] { const char* s = getenv("REGAT_TC_SPLITS"); return s ? atoi(s) : 0; }();
    if (force_splits > 0) splits = force_splits;
    else if (pair_splits > 0) splits = pair_splits;
    else if (split_k > 1) splits = split_k;
    else if (auto_split && tiles_m * tiles_n < num_sms()) splits = (2 * num_sms()) / (tiles_m * tiles_n);
    splits = std::max(1, std::min(splits, total_kb / 8));
  }
  const int kbps = ceil_div(total_kb, splits);
  splits = ceil_div(total_kb, kbps);

  if (!a_mn) REGAT_TRY(make_map(&out.ma, A, K, M, lda, BM)); else REGAT_TRY(make_map(&out.ma, A, M, K, lda, BK));
  if (!b_mn) REGAT_TRY(make_map(&out.mb, B, K, N, ldb, cta2 ? bn / 2 : bn)); else REGAT_TRY(make_map(&out.mb, B, N, K, ldb, BK));

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
  p.trace = g_trace;
  static const int dbg_env = [] { const char* s = getenv("REGAT_TC_DBG"); return s ? atoi(s) : 0; }();
  p.dbg = dbg_env;
  out.a_mn = a_mn; out.b_mn = b_mn; out.bn = bn; out.cta2 = cta2;
  return REGAT_OK;
}

int launch_prepared(const Prepared& x, const Prepared* y, cudaStream_t st, int max_ctas = 0) {
  TcParams p = x.p;
  if (y) {
    p.s2.units = y->p.tiles_m * y->p.tiles_n; p.s2.M = y->p.M; p.s2.tiles_m = y->p.tiles_m; p.s2.total_k_blocks = y->p.total_k_blocks;
    p.s2.accumulate = y->p.e.accumulate; p.s2.C = y->p.C;
  }
  const int units = p.tiles_m * p.tiles_n * p.splits + p.s2.units;
  // BN=256: 4 stages x 48 KB + 2 x 256 TMEM columns, one CTA per SM.  BN=128: 3 stages x 32 KB + 2 x 128 columns, two per SM.
  // REGAT_SM_RESERVE=n leaves n SMs free of persistent GEMM CTAs so that a concurrent NCCL all-reduce can make progress
  static const int reserve = [] { const char* s = getenv("REGAT_SM_RESERVE"); return s ? std::max(0, atoi(s)) : 0; }();
  // max_ctas > 0: a launch that must not take every SM (side-stream products beside the main stream's dependency chain)
  const int sms = std::max(2, std::min(num_sms() - reserve, max_ctas > 0 ? max_ctas : num_sms()));
  const CUtensorMap& ma2 = y ? y->ma : x.ma;
  const CUtensorMap& mb2 = y ? y->mb : x.mb;
  const int epi = epi_kind(p, y != nullptr, y ? y->p.e.accumulate : 0, y ? y->p.C : nullptr);
  // CTA pairs: 6 stages x 32 KB per CTA, grid = an even number of CTAs, two per unit in flight
  if (x.bn == 256 && x.cta2) return launch_cfg<256, 6, 2, 8, true>(x.a_mn, x.b_mn, epi, x.ma, x.mb, p, ma2, mb2, 2 * std::min(units, sms / 2), st);
  if (x.bn == 256) return launch_cfg<256, 4, 2, 8>(x.a_mn, x.b_mn, epi, x.ma, x.mb, p, ma2, mb2, std::min(units, sms), st);
  if (x.bn == 64) return launch_cfg<64, 8, 2, 4>(x.a_mn, x.b_mn, epi, x.ma, x.mb, p, ma2, mb2, std::min(units, sms), st);
  return launch_cfg<128, 3, 2, 4>(x.a_mn, x.b_mn, epi, x.ma, x.mb, p, ma2, mb2, std::min(units, 2 * sms), st);
}

}  // namespace

void gemm_tc_set_trace(long long* buf) { g_trace = buf; }

bool gemm_tc_supported(int transA, int transB, int M, int N, int K, const void* A, int lda, const void* B, int ldb) {
  (void)transA; (void)transB; (void)M; (void)N; (void)K;
  return aligned16(A) && aligned16(B) && lda % 8 == 0 && ldb % 8 == 0;
}

int gemm_tc(int transA, int transB, int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C, int ldc,
            int c_dtype, const EpiArgs& e, int split_k, cudaStream_t st, int c_block_cols, const long long* c_block_off, int max_ctas) {
  if (M <= 0 || N <= 0) return REGAT_OK;
  Prepared x;
  REGAT_TRY(prepare(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, c_dtype, e, split_k, st, c_block_cols, c_block_off, 0, x));
  return launch_prepared(x, nullptr, st, max_ctas);
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
    ok = x.p.splits == 1 && y.p.splits == 1 && x.bn == 256 && x.bn == y.bn && x.cta2 == y.cta2 && x.p.tiles_n == y.p.tiles_n;
  }
  if (!ok) { REGAT_TRY(single(c0)); return single(c1); }
  return launch_prepared(x, &y, st);
}


// ---- chain launches (see gemm_chain_kernel)
struct ChainBuilder::Impl { ChainParams cp; int max_units; bool ok; };
ChainBuilder::ChainBuilder(unsigned int* counter) : impl(new Impl) {
  memset(&impl->cp, 0, sizeof(impl->cp));
  impl->cp.counter = counter; impl->max_units = 0; impl->ok = true;
}
ChainBuilder::~ChainBuilder() { delete impl; }
int ChainBuilder::product(int transA, int transB, int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C, int ldc,
                          int c_dtype, const ChainEpi& e) {
  ChainParams& cp = impl->cp;
  REGAT_REQUIRE(cp.nprob < CH_MAXP, REGAT_ERR_UNSUPPORTED, "chain: too many stages");
  REGAT_REQUIRE(gemm_tc_supported(transA, transB, M, N, K, A, lda, B, ldb), REGAT_ERR_ALIGN, "chain: unaligned operands");
  const int es = c_dtype == REGAT_F32 ? 4 : 2;
  REGAT_REQUIRE(aligned16(C) && (ldc * es) % 16 == 0, REGAT_ERR_ALIGN, "chain: unaligned output");
  REGAT_REQUIRE(c_dtype == REGAT_F32 || N % 32 == 0, REGAT_ERR_UNSUPPORTED, "chain: bf16 outputs need N %% 32 == 0");
  REGAT_REQUIRE((!e.gate && !e.mul && !e.out2) || N % 32 == 0, REGAT_ERR_UNSUPPORTED, "chain: gate / mul terms need N %% 32 == 0");
  REGAT_REQUIRE((!e.gate || (aligned16(e.gate) && e.gate_ld % 8 == 0)) && (!e.mul || (aligned16(e.mul) && e.mul_ld % 8 == 0)) &&
                    (!e.out2 || (aligned16(e.out2) && e.out2_ld % 8 == 0 && e.mul2 && aligned16(e.mul2) && e.mul2_ld % 8 == 0)),
                REGAT_ERR_ALIGN, "chain: unaligned epilogue tensors");
  const int pi = cp.nprob++;
  ChainProblem& P = cp.prob[pi];
  memset(&P, 0, sizeof(P));
  const bool a_mn = transA != 0, b_mn = transB == 0;
  if (!a_mn) REGAT_TRY(make_map(&cp.map[2 * pi], A, K, M, lda, BM)); else REGAT_TRY(make_map(&cp.map[2 * pi], A, M, K, lda, BK));
  if (!b_mn) REGAT_TRY(make_map(&cp.map[2 * pi + 1], B, K, N, ldb, CH_BN)); else REGAT_TRY(make_map(&cp.map[2 * pi + 1], B, N, K, ldb, BK));
  P.kind = 0; P.M = M; P.N = N; P.total_kb = ceil_div(K, BK); P.tiles_n = ceil_div(N, CH_BN); P.units = ceil_div(M, BM) * P.tiles_n;
  P.a_mn = a_mn; P.b_mn = b_mn; P.C = C; P.ldc = ldc; P.c_f32 = c_dtype == REGAT_F32;
  P.bias = e.bias; P.relu = e.relu; P.gate = static_cast<const bf16*>(e.gate); P.gate_ld = e.gate_ld;
  P.mul = static_cast<const bf16*>(e.mul); P.mul_ld = e.mul_ld; P.out2 = static_cast<bf16*>(e.out2); P.out2_ld = e.out2_ld;
  P.mul2 = static_cast<const bf16*>(e.mul2); P.mul2_ld = e.mul2_ld;
  impl->max_units = std::max(impl->max_units, P.units);
  return REGAT_OK;
}
int ChainBuilder::loss(int B, int A, const float* logits, int ldl, const float* target, float gscale, float* loss, float* score, void* dlog,
                       int ldd) {
  ChainParams& cp = impl->cp;
  REGAT_REQUIRE(cp.nprob < CH_MAXP, REGAT_ERR_UNSUPPORTED, "chain: too many stages");
  ChainProblem& P = cp.prob[cp.nprob++];
  memset(&P, 0, sizeof(P));
  P.kind = 1; P.M = B; P.logits = logits; P.ldl = ldl; P.target = target; P.A = A; P.inv_B = 1.f / (float)B; P.gscale = gscale;
  P.loss = loss; P.score = score; P.dlog = static_cast<bf16*>(dlog); P.ldd = ldd;
  impl->max_units = std::max(impl->max_units, std::min(B, num_sms()));   // one row per CTA and round: worth every SM
  return REGAT_OK;
}
bool chain_fits(int max_units) { return max_units <= 2 * num_sms(); }
int ChainBuilder::launch(cudaStream_t st) {
  ChainParams& cp = impl->cp;
  if (cp.nprob == 0) return REGAT_OK;
  constexpr size_t smem = (size_t)CH_STAGES * (BM * BK * 2 + CH_BN * BK * 2) + (2 * CH_STAGES + 2 * CH_ACC) * 8 + 16 + 64 + CH_EPIW * 256;
  static std::mutex mu;
  static bool attr_set[64] = {};
  int dev = 0;
  REGAT_CUDA(cudaGetDevice(&dev));
  {
    std::lock_guard<std::mutex> lk(mu);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
      REGAT_CUDA(cudaFuncSetAttribute(gemm_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
  }
  cp.trace = g_trace;
  // every CTA waits for every other one between stages: the grid must be resident as a whole (one CTA per SM)
  const int grid = std::max(1, std::min(impl->max_units, num_sms()));
  gemm_chain_kernel<<<grid, 64 + 32 * CH_EPIW, smem, st>>>(cp);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}

}  // namespace regat
