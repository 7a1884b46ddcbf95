// Kernel (a): fused box-geometry bias + multi-head graph attention, forward and backward.
//
// What it replaces in the reference (file:line under the reference tree):
//   position_emb.py:117-151  pairwise log-geometry            \  recomputed on chip per pair,
//   position_emb.py:96-115   sinusoidal embedding (64-d)      /  never written to HBM
//   graph_att_layer.py:74-88 pair_pos_fc 64->H, relu, max(.,1e-6), log, raw reshape scramble
//   graph_att_layer.py:63-68 QK^T/sqrt(dh);  :90-102 adjacency where (no-op, adj==1) + label const
//   graph_att_layer.py:104-117 softmax over the first M objects, aggregation, grouped 1x1 conv
//                            (re-associated: V' = s[:, :M].Kc + bc is an input here)
//   graph_att_net.py:64-81   sum over directions, ReLU;  relation_encoder.py:88-91 residual
//
// Design (B200): the path is HBM-bound (~3 FLOP/B), so everything is sized for bytes:
//   * forward: one CTA per (graph, 16-row tile).  Phase 1: 256 threads compute the 16 x M pair
//     embeddings (32 accurate sincosf each) and the 64 x (dirs*H) projection once for ALL heads and
//     both directions, leaving log-bias in shared memory.  Phase 2: warp w owns heads w, w+8,...;
//     Q/K/V' fragments are loaded straight from global with 128-bit loads (k- and n-index
//     permutations make every mma fragment a contiguous load, so there is no smem staging and
//     no bank conflict), logits/softmax live in mma C fragments, P feeds the PV mma with no
//     shuffle, and the epilogue fuses  v1 = v0 + relu(s + O_0 + O_1).
//   * the tiny matmuls run on mma.sync.m16n8k8 tf32.  bf16 inputs are exact in tf32; in fp32 parity
//     mode every product is done as 3xTF32 (hi*hi + hi*lo + lo*hi), which is fp32-accurate.
//   * backward: one CTA per (graph, dir, head): phase A (row tiles over warps) dP, dL, dQ; phase B
//     (head-dim slices over warps) dK, dV' from shared-memory copies; dL overwrites P in place and a
//     separate pair-tiled kernel reduces dW_g/db_g with recomputed embeddings.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace regat {
namespace {

constexpr int EMB = 64;        // imp_pos_emb_dim (butd_vqa.json:11); 8 wavelengths x {sin,cos} x 4 terms
constexpr int HD = 64;         // head dim (graph_att_layer.py:23 with 1024/16)
constexpr int ROWS = 16;       // query rows per forward CTA = one mma m-tile
constexpr int MAX_ROIS = 128;
constexpr int MAX_DH = 32;      // dir_num * num_heads handled by one CTA

struct WaveDiv { float d[8]; };

// ------------------------------------------------------------------------------------------
// geometry (same fp32 operation order as position_emb.py:122-142 and :104-108)
__device__ __forceinline__ float4 box_terms(const float* bx) {
  float x1 = bx[0], y1 = bx[1], x2 = bx[2], y2 = bx[3];
  return make_float4(x2 - x1 + 1.f, y2 - y1 + 1.f, 0.5f * (x1 + x2), 0.5f * (y1 + y2));  // w,h,cx,cy
}

__device__ __forceinline__ void pair_log_geometry(const float4& oi, const float4& oj, float (&P)[4]) {
  float dx = fabsf(__fdiv_rn(oi.z - oj.z, oi.x));
  float dy = fabsf(__fdiv_rn(oi.w - oj.w, oi.y));
  P[0] = logf(dx < 1e-3f ? 1e-3f : dx);
  P[1] = logf(dy < 1e-3f ? 1e-3f : dy);
  P[2] = logf(__fdiv_rn(oi.x, oj.x));
  P[3] = logf(__fdiv_rn(oi.y, oj.y));
}

// sin and cos for |x| < ~1e4 (here |x| <= 100*|log 1e-3| = 691): two-constant Cody-Waite reduction to [-pi/4, pi/4] and
// degree-7/8 minimax polynomials; absolute error ~1e-7, i.e. the accuracy of sincosf() without its slow-path machinery
// (the library call was ~85 instructions per evaluation in this kernel).
__device__ __forceinline__ void sincos_cw(float x, float* sn, float* cs) {
  const float n = rintf(x * 0.636619772367581f);            // x * 2/pi
  const int q = (int)n;
  float r = fmaf(n, -1.57079637050628662f, x);               // pi/2 = hi + lo,  hi = float(pi/2)
  r = fmaf(n, 4.37113882867379e-08f, r);                     // -lo
  const float z = r * r;
  float sp = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
  sp = fmaf(sp, z, -1.6666654611e-1f);
  sp = fmaf(sp * z, r, r);
  float cp = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
  cp = fmaf(cp, z, 4.166664568298827e-2f);
  cp = fmaf(cp, z * z, fmaf(z, -0.5f, 1.0f));
  const float s1 = (q & 1) ? cp : sp, c1 = (q & 1) ? sp : cp;
  *sn = (q & 2) ? -s1 : s1;
  *cs = ((q + 1) & 2) ? -c1 : c1;
}

// one term c of the pairwise log-geometry (position_emb.py:127-142), branch-free
__device__ __forceinline__ float pair_log_term(const float4& oi, const float4& oj, int c) {
  const float num = c == 0 ? oi.z - oj.z : (c == 1 ? oi.w - oj.w : (c == 2 ? oi.x : oi.y));
  const float den = c == 0 ? oi.x : (c == 1 ? oi.y : (c == 2 ? oj.x : oj.y));
  float qv = __fdiv_rn(num, den);
  if (c < 2) { qv = fabsf(qv); qv = qv < 1e-3f ? 1e-3f : qv; }
  return logf(qv);
}

// feature c*16+k = sin(100*P_c/div_k), c*16+8+k = cos(...)   (position_emb.py:104-113)
__device__ __forceinline__ void pair_embedding(const float4& oi, const float4& oj, const WaveDiv& wd,
                                               float (&emb)[EMB]) {
  float P[4];
  pair_log_geometry(oi, oj, P);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float x = 100.0f * P[c];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float sn, cs;
      sincos_cw(__fdiv_rn(x, wd.d[k]), &sn, &cs);
      emb[c * 16 + k] = sn;
      emb[c * 16 + 8 + k] = cs;
    }
  }
}

// ------------------------------------------------------------------------------------------
// mma.sync m16n8k8 tf32 helpers
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// Operand prepared once, used for 1 (bf16 mode) or 3 (fp32 mode, 3xTF32) mma.
// cvt.rna.tf32.f32 is emulated with ~6 integer instructions on sm_100a (measured: it was 40% of this kernel), so:
//   * single-pass mode feeds the fp32 bit pattern as is -- the tensor core reads only the upper 19 bits (truncation; bf16
//     inputs are exact, probabilities lose < 2^-10 relative, inside the bf16-mode tolerance);
//   * 3xTF32 splits with one mask: hi = x & ~0x1fff (exact tf32), lo = x - hi (exact in fp32, then truncated by the MMA):
//     hi*hi + hi*lo + lo*hi reproduces the fp32 product to ~2^-21.
template <bool S3, int NR> struct Opnd {
  uint32_t hi[NR];
  uint32_t lo[S3 ? NR : 1];
  __device__ __forceinline__ void set(int i, float x) {
    if constexpr (S3) {
      hi[i] = __float_as_uint(x) & 0xffffe000u;
      lo[i] = __float_as_uint(x - __uint_as_float(hi[i]));
    } else {
      hi[i] = __float_as_uint(x);
    }
  }
  template <bool EXACT> __device__ __forceinline__ void put(int i, float x) { set(i, x); }
};
template <bool S3>
__device__ __forceinline__ void mma_acc(float (&c)[4], const Opnd<S3, 4>& a, const Opnd<S3, 2>& b) {
  if constexpr (S3) {
    mma_tf32(c, a.lo, b.hi);
    mma_tf32(c, a.hi, b.lo);
  }
  mma_tf32(c, a.hi, b.hi);
}

// ------------------------------------------------------------------------------------------
// fragment loads.  A 64-wide head slice of one row is split over the 4 lanes of a quad (t = lane%4) so
// that every lane issues only 128-bit loads; k-step s (0..7) of the mma takes (v[2s], v[2s+1]) as its
// (k=t, k=t+4) pair.  A and B rows use the same map, so the dot product is unchanged.
template <typename T> __device__ __forceinline__ void load_row16(const T* row, int t, float (&v)[16]);
template <> __device__ __forceinline__ void load_row16<float>(const float* row, int t, float (&v)[16]) {
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    float4 x = __ldg(reinterpret_cast<const float4*>(row + 16 * kk + 4 * t));
    v[4 * kk] = x.x; v[4 * kk + 1] = x.y; v[4 * kk + 2] = x.z; v[4 * kk + 3] = x.w;
  }
}
template <> __device__ __forceinline__ void load_row16<bf16>(const bf16* row, int t, float (&v)[16]) {
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    uint4 x = __ldg(reinterpret_cast<const uint4*>(row + 32 * kk + 8 * t));
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[8 * kk + 2 * i] = __uint_as_float(w[i] << 16);
      v[8 * kk + 2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
}
// which element (0..63) of the head slice v[idx] is
template <typename T> __device__ __forceinline__ int row16_elem(int idx, int t) {
  if constexpr (sizeof(T) == 4) return 16 * (idx >> 2) + 4 * t + (idx & 3);
  else return 32 * (idx >> 3) + 8 * t + (idx & 7);
}
// 8 contiguous elements
template <typename T> __device__ __forceinline__ void load8c(const T* p, float (&v)[8]);
template <> __device__ __forceinline__ void load8c<float>(const float* p, float (&v)[8]) {
  float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + 4));
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <> __device__ __forceinline__ void load8c<bf16>(const bf16* p, float (&v)[8]) {
  uint4 x = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <typename T> __device__ __forceinline__ void store_n(T* p, const float* v, int n);  // n multiple of 4 (fp32) / 8 (bf16)
template <> __device__ __forceinline__ void store_n<float>(float* p, const float* v, int n) {
#pragma unroll
  for (int i = 0; i < n; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}
template <> __device__ __forceinline__ void store_n<bf16>(bf16* p, const float* v, int n) {
#pragma unroll
  for (int i = 0; i < n; i += 8) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[i + 2 * k], v[i + 2 * k + 1]);
      w[k] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p + i) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
template <typename T> __device__ __forceinline__ void store4(T* p, float a, float b, float c, float d) {
  if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
  } else {
    __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
    *reinterpret_cast<uint2*>(p) = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
  }
}

__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}

// padded key stride of the per-CTA bias tile: multiple of 8 and == 8 (mod 16) so that the 64-bit
// C-fragment reads of a half-warp hit 16 distinct bank pairs.
__host__ __device__ inline int bias_stride(int M) {
  int mp = (M + 7) / 8 * 8;
  return (mp % 16 == 0) ? mp + 8 : mp;
}

// ==========================================================================================
// forward
struct FwdParams {
  int B, N, M, D, H, dirs, MP, residual;
  const void* q; const void* kv;
  const float* boxes; const float* pos_emb; WaveDiv wd;
  const float* wg; long long wg_stride; const float* alpha_g; const float* bg; long long bg_stride;
  const float* label_c;
  const void* s; const void* v0; void* v1;
  float* save_p; float* save_gb; unsigned long long* gate;
  // explicit relation (graph_att_layer.py:90-102): per-pair additive term [B][dirs][N][M] = label bias where the adjacency is set,
  // -9e15 where it is not (in fp32 that absorbs the affinity and the label bias exactly as where(adj > 0, aff, -9e15) + label does);
  // replaces the geometry bias (pos_emb_dim = -1: the layer has no pair_pos_fc)
  const float* pair_bias;
};

template <typename T, bool S3, int NTS>
__global__ void __launch_bounds__(256, 2) geoattn_fwd_kernel(const FwdParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int DH = p.dirs * p.H;
  float4* obj = reinterpret_cast<float4*>(smem_raw);                 // [MAX_ROIS]
  float* wgs = reinterpret_cast<float*>(obj + MAX_ROIS);             // [EMB][DH]
  float* bgs = wgs + EMB * DH;                                       // [DH]
  float* ags = bgs + DH;                                             // [DH] alpha per (d,h)
  float* tile = ags + DH;                                            // [DH][ROWS][MP]
  const int tid = threadIdx.x, b = blockIdx.y, i0 = blockIdx.x * ROWS;
  const int N = p.N, M = p.M, MP = p.MP, D = p.D, H = p.H;

  if (p.pair_bias) {
    // explicit relation: the additive term is an input, identical for all heads of a direction
    for (int x = tid; x < DH * ROWS * M; x += 256) {
      const int dh = x / (ROWS * M), rem = x - dh * (ROWS * M), il = rem / M, j = rem - il * M, d = dh / H;
      const int i = min(i0 + il, N - 1);
      tile[(dh * ROWS + il) * MP + j] = __ldg(p.pair_bias + (((size_t)b * p.dirs + d) * N + i) * M + j);
    }
  } else {
  // ---- phase 0: per-object terms and the pair_pos_fc weights into shared memory
  for (int n = tid; n < N; n += 256) obj[n] = p.boxes ? box_terms(p.boxes + ((size_t)b * N + n) * 4) : make_float4(1, 1, 0, 0);
  // pair_pos_fc kernels of all directions, stored in mma B-fragment order: wgs[((ks*NTD + nt)*32 + lane)*2 + r] =
  // W_g[feature 8ks + lane%4 + 4r][dh = 8nt + lane/4]  -> every lane reads its (b0, b1) with one conflict-free 64-bit load
  const int NTD = DH >> 3;
  for (int x = tid; x < EMB * DH; x += 256) {
    const int r = x & 1, ln = (x >> 1) & 31, rest = x >> 6, nt = rest % NTD, ks = rest / NTD;
    const int e = 8 * ks + (ln & 3) + 4 * r, dh = 8 * nt + (ln >> 2), d = dh / H, h = dh - d * H;
    wgs[x] = p.wg[(size_t)d * p.wg_stride + e * H + h];
  }
  for (int dh = tid; dh < DH; dh += 256) {
    int d = dh / H, h = dh - d * H;
    bgs[dh] = p.bg ? p.bg[(size_t)d * p.bg_stride + h] : 0.f;
    ags[dh] = p.alpha_g[d];
  }
  __syncthreads();

  // ---- phase 1: geometry bias for the 16 x M pairs of this row tile, all heads and directions.
  // bias[i,j] belongs to box pair (f / N, f % N), f = i*M + j  (graph_att_layer.py:74,81 raw reshape of the row-sliced [M,N]
  // tensor from position_emb.py:146).  The 64 -> dirs*H projection runs on the tensor cores: a warp takes 16 consecutive
  // pairs as one mma m-tile; lane (g, t) evaluates exactly the sin/cos values its A fragments need (pairs g and g+8, wave
  // numbers t and t+4 of every geometry term -- 16 sincos per lane per tile, none computed twice), the quad shares the four
  // log-geometry terms by shuffle, and z lands in C fragments where alpha, bias, relu, max and log are applied.
  {
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int npairs = ROWS * M;
    for (int mt = warp; mt * 16 < npairs; mt += 8) {
      int il[2], jj[2], fidx[2];
      bool inb[2], rowok[2];
      float mine[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int pi = mt * 16 + g + 8 * u;
        inb[u] = pi < npairs;
        il[u] = min(pi, npairs - 1) / M; jj[u] = min(pi, npairs - 1) - il[u] * M;
        const int i = i0 + il[u];
        rowok[u] = inb[u] && i < N;
        fidx[u] = min(i, N - 1) * M + jj[u];
        const int ip = fidx[u] / N, jp = fidx[u] - ip * N;
        mine[u] = p.pos_emb ? 0.f : pair_log_term(obj[ip], obj[jp], t);     // lane t owns geometry term t of its two pairs
      }
      float zc[MAX_DH / 8][4];
#pragma unroll
      for (int nt = 0; nt < MAX_DH / 8; ++nt) { zc[nt][0] = zc[nt][1] = zc[nt][2] = zc[nt][3] = 0.f; }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float sv[2][2], cv[2][2];        // [pair][wave number t / t+4]
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (p.pos_emb) {
            const float* er = p.pos_emb + ((size_t)b * M * N + fidx[u]) * EMB + c * 16;
            sv[u][0] = __ldg(er + t); sv[u][1] = __ldg(er + t + 4); cv[u][0] = __ldg(er + 8 + t); cv[u][1] = __ldg(er + 12 + t);
          } else {
            const float x = 100.0f * __shfl_sync(0xffffffffu, mine[u], (lane & ~3) | c);
            sincos_cw(__fdiv_rn(x, p.wd.d[t]), &sv[u][0], &cv[u][0]);
            sincos_cw(__fdiv_rn(x, p.wd.d[t + 4]), &sv[u][1], &cv[u][1]);
          }
        }
        Opnd<true, 4> as, ac;      // 3xTF32 in both modes: log() amplifies any error of z near 0
        as.set(0, sv[0][0]); as.set(1, sv[1][0]); as.set(2, sv[0][1]); as.set(3, sv[1][1]);
        ac.set(0, cv[0][0]); ac.set(1, cv[1][0]); ac.set(2, cv[0][1]); ac.set(3, cv[1][1]);
#pragma unroll
        for (int nt = 0; nt < MAX_DH / 8; ++nt) {
          if (nt < NTD) {
            const float2 ws = *reinterpret_cast<const float2*>(wgs + (((2 * c) * NTD + nt) * 32 + lane) * 2);
            const float2 wc = *reinterpret_cast<const float2*>(wgs + (((2 * c + 1) * NTD + nt) * 32 + lane) * 2);
            Opnd<true, 2> bs, bc;
            bs.set(0, ws.x); bs.set(1, ws.y); bc.set(0, wc.x); bc.set(1, wc.y);
            mma_acc<true>(zc[nt], as, bs);
            mma_acc<true>(zc[nt], ac, bc);
          }
        }
      }
#pragma unroll
      for (int nt = 0; nt < MAX_DH / 8; ++nt) {
        if (nt < NTD) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int u = e >> 1, dh = 8 * nt + 2 * t + (e & 1);
            if (inb[u]) {
              float gb = 0.f;
              if (rowok[u]) {
                const float z = fmaf(ags[dh], zc[nt][e], bgs[dh]);
                gb = logf(fmaxf(fmaxf(z, 0.f), 1e-6f));                   // graph_att_layer.py:79,86,88
                if (p.save_gb) p.save_gb[(((size_t)b * DH + dh) * N + i0 + il[u]) * M + jj[u]] = gb;
              }
              tile[(dh * ROWS + il[u]) * MP + jj[u]] = gb;
            }
          }
        }
      }
    }
  }
  }  // geometry bias
  __syncthreads();

  // ---- phase 2: attention.  warp <-> head; lane (g = lane/4, t = lane%4) in mma fragment terms.
  constexpr bool EX = sizeof(T) == 2;      // bf16 inputs are exact tf32 values
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const T* Q = static_cast<const T*>(p.q);
  const T* KV = static_cast<const T*>(p.kv);
  const int ldq = p.dirs * D, ldkv = 2 * p.dirs * D;
  const float c_label = p.label_c ? __ldg(p.label_c) : 0.f;
  const int r0 = i0 + g, r1 = i0 + g + 8;                      // the two query rows this lane touches
  const int r0c = min(r0, N - 1), r1c = min(r1, N - 1);

  for (int h = warp; h < H; h += 8) {
    float oacc[8][4];
#pragma unroll
    for (int ot = 0; ot < 8; ++ot) { oacc[ot][0] = oacc[ot][1] = oacc[ot][2] = oacc[ot][3] = 0.f; }

    for (int d = 0; d < p.dirs; ++d) {
      const int dh = d * H + h;
      // S = Q K^T
      float sacc[NTS][4];
#pragma unroll
      for (int nt = 0; nt < NTS; ++nt) { sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f; }
      constexpr bool PACKED = sizeof(T) == 2 && !S3;
      // bf16 mode: every operand of this (dir, head) is requested up front as packed 128-bit registers -- one memory round
      // trip instead of one per key tile -- and unpacked (a shift / a mask per value) right at the mma.
      uint4 qw[2][2], kw[NTS][2], vw[NTS][2];
      if constexpr (PACKED) {
        const bf16* qp0 = reinterpret_cast<const bf16*>(Q) + ((size_t)b * N + r0c) * ldq + d * D + h * HD + 8 * t;
        const bf16* qp1 = reinterpret_cast<const bf16*>(Q) + ((size_t)b * N + r1c) * ldq + d * D + h * HD + 8 * t;
#pragma unroll
        for (int kk = 0; kk < 2; ++kk) {
          qw[0][kk] = __ldg(reinterpret_cast<const uint4*>(qp0 + 32 * kk));
          qw[1][kk] = __ldg(reinterpret_cast<const uint4*>(qp1 + 32 * kk));
        }
#pragma unroll
        for (int nt = 0; nt < NTS; ++nt) {
          if (nt * 8 < M) {
            const bf16* kp = reinterpret_cast<const bf16*>(KV) + ((size_t)b * M + min(nt * 8 + g, M - 1)) * ldkv + d * D + h * HD + 8 * t;
            kw[nt][0] = __ldg(reinterpret_cast<const uint4*>(kp));
            kw[nt][1] = __ldg(reinterpret_cast<const uint4*>(kp + 32));
            const int j0 = min(nt * 8 + 2 * t, M - 1), j1 = min(nt * 8 + 2 * t + 1, M - 1);
            const bf16* vp = reinterpret_cast<const bf16*>(KV) + (size_t)b * M * ldkv + (p.dirs + d) * D + h * HD + 8 * g;
            vw[nt][0] = __ldg(reinterpret_cast<const uint4*>(vp + (size_t)j0 * ldkv));
            vw[nt][1] = __ldg(reinterpret_cast<const uint4*>(vp + (size_t)j1 * ldkv));
          }
        }
#pragma unroll
        for (int nt = 0; nt < NTS; ++nt) {
          if (nt * 8 < M) {
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const uint32_t q0[4] = {qw[0][kk].x, qw[0][kk].y, qw[0][kk].z, qw[0][kk].w};
              const uint32_t q1[4] = {qw[1][kk].x, qw[1][kk].y, qw[1][kk].z, qw[1][kk].w};
              const uint32_t k4[4] = {kw[nt][kk].x, kw[nt][kk].y, kw[nt][kk].z, kw[nt][kk].w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint32_t a[4] = {q0[i] << 16, q1[i] << 16, q0[i] & 0xffff0000u, q1[i] & 0xffff0000u};
                const uint32_t bb[2] = {k4[i] << 16, k4[i] & 0xffff0000u};
                mma_tf32(sacc[nt], a, bb);
              }
            }
          }
        }
      } else {
        float qa[16], qb[16];
        load_row16<T>(Q + ((size_t)b * N + r0c) * ldq + d * D + h * HD, t, qa);
        load_row16<T>(Q + ((size_t)b * N + r1c) * ldq + d * D + h * HD, t, qb);
#pragma unroll
        for (int nt = 0; nt < NTS; ++nt) {
          if (nt * 8 < M) {
            float kr[16];
            load_row16<T>(KV + ((size_t)b * M + min(nt * 8 + g, M - 1)) * ldkv + d * D + h * HD, t, kr);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              Opnd<S3, 4> a; Opnd<S3, 2> bb;
              a.template put<EX>(0, qa[2 * ks]); a.template put<EX>(1, qb[2 * ks]);
              a.template put<EX>(2, qa[2 * ks + 1]); a.template put<EX>(3, qb[2 * ks + 1]);
              bb.template put<EX>(0, kr[2 * ks]); bb.template put<EX>(1, kr[2 * ks + 1]);
              mma_acc<S3>(sacc[nt], a, bb);
            }
          }
        }
      }
      // logits = S/sqrt(dh) + geometry bias + label const; mask padded key columns; softmax over keys
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NTS; ++nt) {
        const int c0 = nt * 8 + 2 * t;
        const float2 b0 = *reinterpret_cast<const float2*>(tile + (dh * ROWS + g) * MP + min(c0, MP - 2));
        const float2 b1 = *reinterpret_cast<const float2*>(tile + (dh * ROWS + g + 8) * MP + min(c0, MP - 2));
        sacc[nt][0] = c0 < M ? fmaf(sacc[nt][0], 0.125f, b0.x + c_label) : -INFINITY;
        sacc[nt][1] = c0 + 1 < M ? fmaf(sacc[nt][1], 0.125f, b0.y + c_label) : -INFINITY;
        sacc[nt][2] = c0 < M ? fmaf(sacc[nt][2], 0.125f, b1.x + c_label) : -INFINITY;
        sacc[nt][3] = c0 + 1 < M ? fmaf(sacc[nt][3], 0.125f, b1.y + c_label) : -INFINITY;
        mx0 = fmaxf(mx0, fmaxf(sacc[nt][0], sacc[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(sacc[nt][2], sacc[nt][3]));
      }
      mx0 = quad_max(mx0); mx1 = quad_max(mx1);
      float sm0 = 0.f, sm1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < NTS; ++nt) {
        sacc[nt][0] = expf(sacc[nt][0] - mx0); sacc[nt][1] = expf(sacc[nt][1] - mx0);
        sacc[nt][2] = expf(sacc[nt][2] - mx1); sacc[nt][3] = expf(sacc[nt][3] - mx1);
        sm0 += sacc[nt][0] + sacc[nt][1]; sm1 += sacc[nt][2] + sacc[nt][3];
      }
      const float inv0 = 1.f / quad_sum(sm0), inv1 = 1.f / quad_sum(sm1);
#pragma unroll
      for (int nt = 0; nt < NTS; ++nt) {
        sacc[nt][0] *= inv0; sacc[nt][1] *= inv0; sacc[nt][2] *= inv1; sacc[nt][3] *= inv1;
        if (p.save_p) {
          const int c0 = nt * 8 + 2 * t;
          float* pr0 = p.save_p + (((size_t)b * DH + dh) * N + r0) * M;
          float* pr1 = p.save_p + (((size_t)b * DH + dh) * N + r1) * M;
          if ((M & 1) == 0) {       // rows start 8-byte aligned: one 64-bit store per row
            if (r0 < N && c0 < M) *reinterpret_cast<float2*>(pr0 + c0) = make_float2(sacc[nt][0], sacc[nt][1]);
            if (r1 < N && c0 < M) *reinterpret_cast<float2*>(pr1 + c0) = make_float2(sacc[nt][2], sacc[nt][3]);
          } else {
            if (r0 < N) { if (c0 < M) pr0[c0] = sacc[nt][0]; if (c0 + 1 < M) pr0[c0 + 1] = sacc[nt][1]; }
            if (r1 < N) { if (c0 < M) pr1[c0] = sacc[nt][2]; if (c0 + 1 < M) pr1[c0 + 1] = sacc[nt][3]; }
          }
        }
      }
      // O += P V'.  k-step nt uses keys (8nt+2t, 8nt+2t+1) as (k=t, k=t+4): the C fragment of P is already
      // the A fragment.  Output n-tile ot, column g stands for head-dim element 8g+ot, so a lane loads 8
      // contiguous V' elements per key and ends up owning 16 contiguous output elements per row.
#pragma unroll
      for (int nt = 0; nt < NTS; ++nt) {
        if (nt * 8 < M) {
          if constexpr (PACKED) {
            const uint32_t a[4] = {__float_as_uint(sacc[nt][0]), __float_as_uint(sacc[nt][2]), __float_as_uint(sacc[nt][1]),
                                   __float_as_uint(sacc[nt][3])};
            const uint32_t v0w[4] = {vw[nt][0].x, vw[nt][0].y, vw[nt][0].z, vw[nt][0].w};
            const uint32_t v1w[4] = {vw[nt][1].x, vw[nt][1].y, vw[nt][1].z, vw[nt][1].w};
#pragma unroll
            for (int ot = 0; ot < 8; ++ot) {
              const uint32_t bb[2] = {(ot & 1) ? (v0w[ot >> 1] & 0xffff0000u) : (v0w[ot >> 1] << 16),
                                      (ot & 1) ? (v1w[ot >> 1] & 0xffff0000u) : (v1w[ot >> 1] << 16)};
              mma_tf32(oacc[ot], a, bb);
            }
          } else {
            float va[8], vb[8];
            const int j0 = min(nt * 8 + 2 * t, M - 1), j1 = min(nt * 8 + 2 * t + 1, M - 1);
            load8c<T>(KV + ((size_t)b * M + j0) * ldkv + (p.dirs + d) * D + h * HD + 8 * g, va);
            load8c<T>(KV + ((size_t)b * M + j1) * ldkv + (p.dirs + d) * D + h * HD + 8 * g, vb);
            Opnd<S3, 4> a;
            a.set(0, sacc[nt][0]); a.set(1, sacc[nt][2]); a.set(2, sacc[nt][1]); a.set(3, sacc[nt][3]);
#pragma unroll
            for (int ot = 0; ot < 8; ++ot) {
              Opnd<S3, 2> bb;
              bb.template put<EX>(0, va[ot]); bb.template put<EX>(1, vb[ot]);
              mma_acc<S3>(oacc[ot], a, bb);
            }
          }
        }
      }
    }  // dirs

    // epilogue: lane owns head-dim elements [16t, 16t+16) of rows r0 and r1
    const T* S = static_cast<const T*>(p.s);
    const T* V0 = static_cast<const T*>(p.v0);
    T* V1 = static_cast<T*>(p.v1);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = half ? r1 : r0;
      unsigned long long bits = 0ull;
      if (r < N) {
        const size_t off = ((size_t)b * N + r) * D + h * HD + 16 * t;
        float sv[16], vv[16], out[16];
        if (S) {
          load8c<T>(S + off, *reinterpret_cast<float(*)[8]>(&sv[0]));
          load8c<T>(S + off + 8, *reinterpret_cast<float(*)[8]>(&sv[8]));
        }
        if (p.residual) {
          load8c<T>(V0 + off, *reinterpret_cast<float(*)[8]>(&vv[0]));
          load8c<T>(V0 + off + 8, *reinterpret_cast<float(*)[8]>(&vv[8]));
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
          const float o = oacc[u & 7][(u >> 3) + 2 * half];      // element 16t+u <- tile (u%8), col 2t + u/8
          if (S) {
            const float x = sv[u] + o;
            if (x > 0.f) bits |= 1ull << (16 * t + u);
            out[u] = (p.residual ? vv[u] : 0.f) + fmaxf(x, 0.f);
          } else {
            out[u] = o;      // raw attention output of one GraphSelfAttentionLayer (graph_att_layer.py:121)
          }
        }
        store_n<T>(V1 + off, out, 16);
      }
      if (p.gate) {
        bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
        bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
        if (t == 0 && r < N) p.gate[((size_t)b * N + r) * H + h] = bits;
      }
    }
  }
}

// ==========================================================================================
// forward, bf16 fast path (boxes given, M <= 8 NTS).  Same CTA shape and phase structure as geoattn_fwd_kernel, sized for
// instruction count (the generic kernel is issue-bound at ~270k warp instructions per graph):
//   * phase 1 only walks the pairs of rows that exist; sin/cos go through a 2-constant Cody-Waite reduction to [-pi, pi]
//     and the SFU (abs. error 2^-21, below the 1-ulp argument noise at |x| ~ 690 that the reference itself has across NumPy
//     builds); divisions by the wavelengths become multiplications; the 64 -> dirs*H projection is one TF32 pass;
//   * phase 2 runs on mma.sync.m16n8k16 bf16: the 128-bit global loads ARE the fragments (k-permutation), so QK^T has no
//     unpacking at all, and P V' needs one PRMT per B register.  P is rounded to bf16 only as an MMA operand; the saved
//     probabilities stay fp32.
// The fp32 parity mode keeps the exact kernel above.
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void sincos_sfu(float x, float* sn, float* cs) {
  const float n = rintf(x * 0.15915494309189535f);          // x / 2pi
  float r = fmaf(n, -6.2831854820251465f, x);                // 2pi = hi + lo, hi = float(2pi)
  r = fmaf(n, 1.7484555e-07f, r);                            // -lo
  *sn = __sinf(r);
  *cs = __cosf(r);
}
__device__ __forceinline__ float pair_log_term_fast(const float4& oi, const float4& oj, int c) {
  const float num = c == 0 ? oi.z - oj.z : (c == 1 ? oi.w - oj.w : (c == 2 ? oi.x : oi.y));
  const float den = c == 0 ? oi.x : (c == 1 ? oi.y : (c == 2 ? oj.x : oj.y));
  float qv = __fdividef(num, den);
  if (c < 2) { qv = fabsf(qv); qv = qv < 1e-3f ? 1e-3f : qv; }
  return __logf(qv);
}
// volatile: keeps the load where the source puts it relative to the (volatile) mma instructions -- the V' quads are
// requested after the QK^T mmas have consumed the Q / K registers, and arrive during the softmax
__device__ __forceinline__ uint4 ldg128(const bf16* p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
// word i (compile-time after unrolling) of a 128-bit register quad, without taking its address
__device__ __forceinline__ uint32_t word(const uint4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

// SFU transcendentals without the denormal fix-up code that __expf / __logf carry when the file is not built with -ftz
__device__ __forceinline__ float ex2_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2_ftz(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_ftz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// exact for 0 <= x < 2^20 and 1 <= d <= 4096: (x + 0.5)/d stays >= 0.5/d away from every integer, far more than the rounding
// of the reciprocal and the product
__device__ __forceinline__ int fdiv_small(int x, float inv_d) { return (int)(((float)x + 0.5f) * inv_d); }

struct FastFwdParams {
  int B, N, M, D, H, dirs, MP /* fp32 row stride of the bias tile in shared memory (even) */, MPAD /* bf16 row stride of P / rz */;
  int CR /* query rows per CTA (multiple of 16, or >= N) */, residual;
  const bf16* q; const bf16* kv; const float* boxes; WaveDiv wd;
  const float* wg; long long wg_stride; const float* alpha_g; const float* bg; long long bg_stride; const float* label_c;
  const bf16* s; const bf16* v0; bf16* v1;
  bf16* save_p; bf16* save_rz; unsigned long long* gate;
};

// One CTA = one graph (all of its query rows when N <= 48, else a chunk of CR rows).
//   phase 1  the log-bias of rows x M pairs for all heads and both directions, once, into a shared-memory tile
//            [dirs*H][rows][MP] fp32.  A warp takes 16 consecutive pairs as one mma m-tile; lane (g, t) evaluates exactly the
//            sin/cos its A fragments need (pairs g, g+8; wave numbers t, t+4), the 64 -> dirs*H projection is one TF32 pass.
//            Index scramble (graph_att_layer.py:74,81) without integer division.
//   phase 2  warp <-> head (h = warp, warp + 8, ...); per (head, 16-row tile, direction): Q K^T on mma.m16n8k16.bf16 (the
//            128-bit global loads ARE the fragments), softmax in the C fragments, P V'.  For training the probabilities leave as
//            the SAME packed bf16 pairs that feed the P V' mma, next to rz = 1/z (0 where the relu / 1e-6 clamp cut the
//            gradient), which the backward kernels consume in the same fragment layout: no fp32 copies of P or of the bias.
//   epilogue v1 = v0 + relu(s + O_0 + O_1) and the 64-bit relu gate per (row, head).
template <int NTS, int DIRS>
__global__ void __launch_bounds__(256, 2) geoattn_fwd_fast_kernel(const FastFwdParams p) {
  constexpr int NKS = (NTS + 1) / 2;             // 16-key k-steps of P V'
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int DH = DIRS * p.H;
  float4* obj = reinterpret_cast<float4*>(smem_raw);                 // [MAX_ROIS]
  float* wgs = reinterpret_cast<float*>(obj + MAX_ROIS);             // [EMB][DH] in B-fragment order
  float* bgs = wgs + EMB * DH;                                       // [DH]
  float* ags = bgs + DH;                                             // [DH]
  float* tile = ags + DH;                                            // [DH][rows][MP]
  const int tid = threadIdx.x, b = blockIdx.y, i0 = blockIdx.x * p.CR;
  const int N = p.N, M = p.M, MP = p.MP, D = p.D, H = p.H;
  const int rows = min(p.CR, N - i0);                                // query rows of this CTA
  const int TS = rows * MP;                                          // floats per (dir, head) slice of the tile
  const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;

  for (int n = tid; n < N; n += 256) obj[n] = box_terms(p.boxes + ((size_t)b * N + n) * 4);
  const int NTD = DH >> 3;
  for (int x = tid; x < EMB * DH; x += 256) {
    const int r = x & 1, ln = (x >> 1) & 31, rest = x >> 6, nt = rest % NTD, ks = rest / NTD;
    const int e = 8 * ks + (ln & 3) + 4 * r, dh = 8 * nt + (ln >> 2), d = dh / H, h = dh - d * H;
    wgs[x] = p.wg[(size_t)d * p.wg_stride + e * H + h];
  }
  for (int dh = tid; dh < DH; dh += 256) {
    int d = dh / H, h = dh - d * H;
    bgs[dh] = p.bg ? p.bg[(size_t)d * p.bg_stride + h] : 0.f;
    ags[dh] = p.alpha_g[d];
  }
  __syncthreads();

  // ---- phase 1
  {
    const int npairs = rows * M;
    const float winv0 = 100.0f / p.wd.d[t], winv1 = 100.0f / p.wd.d[t + 4];
    const float invM = 1.0f / (float)M, invN = 1.0f / (float)N;
    for (int mt = warp; mt * 16 < npairs; mt += 8) {
      int toff[2];
      bool inb[2];
      float mine[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int pi = mt * 16 + g + 8 * u;
        inb[u] = pi < npairs;
        const int pc = min(pi, npairs - 1);
        const int il = fdiv_small(pc, invM), jj = pc - il * M;
        toff[u] = il * MP + jj;
        const int f = i0 * M + pc;                            // = (i0 + il) * M + jj: raw-reshape scramble (graph_att_layer.py:74,81)
        const int ip = fdiv_small(f, invN), jp = f - ip * N;
        mine[u] = pair_log_term_fast(obj[ip], obj[jp], t);
      }
      float zc[MAX_DH / 8][4];
#pragma unroll
      for (int nt = 0; nt < MAX_DH / 8; ++nt) { zc[nt][0] = zc[nt][1] = zc[nt][2] = zc[nt][3] = 0.f; }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float x0 = __shfl_sync(0xffffffffu, mine[0], (lane & ~3) | c);
        const float x1 = __shfl_sync(0xffffffffu, mine[1], (lane & ~3) | c);
        float s00, c00, s01, c01, s10, c10, s11, c11;            // [pair][wave number t / t+4]
        sincos_sfu(x0 * winv0, &s00, &c00); sincos_sfu(x0 * winv1, &s01, &c01);
        sincos_sfu(x1 * winv0, &s10, &c10); sincos_sfu(x1 * winv1, &s11, &c11);
        const uint32_t as[4] = {__float_as_uint(s00), __float_as_uint(s10), __float_as_uint(s01), __float_as_uint(s11)};
        const uint32_t ac[4] = {__float_as_uint(c00), __float_as_uint(c10), __float_as_uint(c01), __float_as_uint(c11)};
#pragma unroll
        for (int nt = 0; nt < MAX_DH / 8; ++nt) {
          if (nt < NTD) {
            const float2 ws = *reinterpret_cast<const float2*>(wgs + (((2 * c) * NTD + nt) * 32 + lane) * 2);
            const float2 wc = *reinterpret_cast<const float2*>(wgs + (((2 * c + 1) * NTD + nt) * 32 + lane) * 2);
            const uint32_t bs[2] = {__float_as_uint(ws.x), __float_as_uint(ws.y)};
            const uint32_t bc[2] = {__float_as_uint(wc.x), __float_as_uint(wc.y)};
            mma_tf32(zc[nt], as, bs);
            mma_tf32(zc[nt], ac, bc);
          }
        }
      }
#pragma unroll
      for (int nt = 0; nt < MAX_DH / 8; ++nt) {
        if (nt < NTD) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int u = e >> 1, dh = 8 * nt + 2 * t + (e & 1);
            if (inb[u]) {
              const float z = fmaf(ags[dh], zc[nt][e], bgs[dh]);
              tile[dh * TS + toff[u]] = 0.6931471806f * lg2_ftz(fmaxf(z, 1e-6f));          // log(max(relu(z), 1e-6)), graph_att_layer.py:79,86,88
            }
          }
        }
      }
    }
  }
  __syncthreads();

  // ---- phase 2
  constexpr float LOG2E = 1.4426950409f;
  const int ldq = DIRS * D, ldkv = 2 * DIRS * D;
  const float c_label = p.label_c ? __ldg(p.label_c) : 0.f;
  const float LOGMIN = -13.815510f;                       // log(1e-6): bias values at the clamp carry no gradient
  const bf16* qg = p.q + (size_t)b * N * ldq + 8 * t;      // + row * ldq + d * D + h * HD (+ 32)
  const bf16* kg = p.kv + (size_t)b * M * ldkv + 8 * t;    // + key * ldkv + d * D + h * HD (+ 32)
  const bf16* vg = p.kv + (size_t)b * M * ldkv + 8 * g;    // + key * ldkv + (dirs + d) * D + h * HD
  // key rows this lane touches (clamped once): K fragments by (nt, g), V' pieces by (ks, t)
  int krow[NTS], vrow[NKS][4];
#pragma unroll
  for (int nt = 0; nt < NTS; ++nt) krow[nt] = min(nt * 8 + g, M - 1) * ldkv;
#pragma unroll
  for (int ks = 0; ks < NKS; ++ks) {
    const int j = 16 * ks + 2 * t;
    vrow[ks][0] = min(j, M - 1) * ldkv; vrow[ks][1] = min(j + 1, M - 1) * ldkv;
    vrow[ks][2] = min(j + 8, M - 1) * ldkv; vrow[ks][3] = min(j + 9, M - 1) * ldkv;
  }
  const int ntiles = (rows + 15) >> 4;

  for (int h = warp; h < H; h += 8) {
    for (int tl = 0; tl < ntiles; ++tl) {
      const int lr0 = tl * 16 + g, lr1 = lr0 + 8;                       // local rows of this lane
      const bool ok0 = lr0 < rows, ok1 = lr1 < rows;
      const int r0 = i0 + min(lr0, rows - 1), r1 = i0 + min(lr1, rows - 1);   // clamped global rows (loads)
      float oacc[8][4];
#pragma unroll
      for (int ot = 0; ot < 8; ++ot) { oacc[ot][0] = oacc[ot][1] = oacc[ot][2] = oacc[ot][3] = 0.f; }

#pragma unroll 1
      for (int d = 0; d < DIRS; ++d) {
        const int dh = d * H + h, col = d * D + h * HD;
        uint4 qw[2][2], kw[NTS][2], vw[NKS][4];
        {
          const bf16* qp0 = qg + (size_t)r0 * ldq + col;
          const bf16* qp1 = qg + (size_t)r1 * ldq + col;
          qw[0][0] = ldg128(qp0); qw[0][1] = ldg128(qp0 + 32);
          qw[1][0] = ldg128(qp1); qw[1][1] = ldg128(qp1 + 32);
          const bf16* kb = kg + col;
#pragma unroll
          for (int nt = 0; nt < NTS; ++nt) { kw[nt][0] = ldg128(kb + krow[nt]); kw[nt][1] = ldg128(kb + krow[nt] + 32); }
        }
        float sacc[NTS][4];
#pragma unroll
        for (int nt = 0; nt < NTS; ++nt) { sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f; }
#pragma unroll
        for (int sst = 0; sst < 4; ++sst) {
          const int w0 = 2 * (sst & 1);
          const uint32_t a[4] = {word(qw[0][sst >> 1], w0), word(qw[1][sst >> 1], w0), word(qw[0][sst >> 1], w0 + 1), word(qw[1][sst >> 1], w0 + 1)};
#pragma unroll
          for (int nt = 0; nt < NTS; ++nt) mma_bf16(sacc[nt], a, word(kw[nt][sst >> 1], w0), word(kw[nt][sst >> 1], w0 + 1));
        }
        {
          const bf16* vb = vg + (DIRS + d) * D + h * HD;
#pragma unroll
          for (int ks = 0; ks < NKS; ++ks) {
            vw[ks][0] = ldg128(vb + vrow[ks][0]); vw[ks][1] = ldg128(vb + vrow[ks][1]);
            vw[ks][2] = ldg128(vb + vrow[ks][2]); vw[ks][3] = ldg128(vb + vrow[ks][3]);
          }
        }
        // logits = S/sqrt(dh) + geometry bias + label const; keys >= M masked; softmax over keys (base-2 exponent)
        const float* trow0 = tile + dh * TS + min(lr0, rows - 1) * MP;
        const float* trow1 = tile + dh * TS + min(lr1, rows - 1) * MP;
        // packed-pair rows of this lane in the saved tensors: [b][dh][row][MPAD] bf16 = 32-bit words [.. * MPAD/2 + t]
        const size_t w0i = (((size_t)b * DH + dh) * N + r0) * (p.MPAD >> 1) + t, w1i = (((size_t)b * DH + dh) * N + r1) * (p.MPAD >> 1) + t;
        uint32_t* rzp0 = reinterpret_cast<uint32_t*>(p.save_rz) + w0i;
        uint32_t* rzp1 = reinterpret_cast<uint32_t*>(p.save_rz) + w1i;
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < NTS; ++nt) {
          const int c0 = nt * 8 + 2 * t;
          const bool v0k = c0 < M, v1k = c0 + 1 < M;
          const int co = v0k ? c0 : 0;                           // column pairs past M re-read pair 0 of the row (values unused)
          const float2 b0 = *reinterpret_cast<const float2*>(trow0 + co);
          const float2 b1 = *reinterpret_cast<const float2*>(trow1 + co);
          if (p.save_rz && v0k) {
            // rz = d log(max(relu(z), 1e-6)) / dz = 1/z above the clamp, 0 at it (graph_att_layer.py:79-88); bias = log z
            const float z0 = b0.x > LOGMIN ? ex2_ftz(-LOG2E * b0.x) : 0.f, z1 = (v1k && b0.y > LOGMIN) ? ex2_ftz(-LOG2E * b0.y) : 0.f;
            const float z2 = b1.x > LOGMIN ? ex2_ftz(-LOG2E * b1.x) : 0.f, z3 = (v1k && b1.y > LOGMIN) ? ex2_ftz(-LOG2E * b1.y) : 0.f;
            if (ok0) rzp0[4 * nt] = pack2_bf16(z0, z1);
            if (ok1) rzp1[4 * nt] = pack2_bf16(z2, z3);
          }
          // logits in base-2 units: (S/8 + bias + c) * log2(e)
          sacc[nt][0] = v0k ? fmaf(sacc[nt][0], 0.125f * LOG2E, (b0.x + c_label) * LOG2E) : -INFINITY;
          sacc[nt][1] = v1k ? fmaf(sacc[nt][1], 0.125f * LOG2E, (b0.y + c_label) * LOG2E) : -INFINITY;
          sacc[nt][2] = v0k ? fmaf(sacc[nt][2], 0.125f * LOG2E, (b1.x + c_label) * LOG2E) : -INFINITY;
          sacc[nt][3] = v1k ? fmaf(sacc[nt][3], 0.125f * LOG2E, (b1.y + c_label) * LOG2E) : -INFINITY;
          mx0 = fmaxf(mx0, fmaxf(sacc[nt][0], sacc[nt][1]));
          mx1 = fmaxf(mx1, fmaxf(sacc[nt][2], sacc[nt][3]));
        }
        mx0 = quad_max(mx0); mx1 = quad_max(mx1);
        float sm0 = 0.f, sm1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < NTS; ++nt) {
          sacc[nt][0] = ex2_ftz(sacc[nt][0] - mx0); sacc[nt][1] = ex2_ftz(sacc[nt][1] - mx0);
          sacc[nt][2] = ex2_ftz(sacc[nt][2] - mx1); sacc[nt][3] = ex2_ftz(sacc[nt][3] - mx1);
          sm0 += sacc[nt][0] + sacc[nt][1]; sm1 += sacc[nt][2] + sacc[nt][3];
        }
        const float inv0 = rcp_ftz(quad_sum(sm0)), inv1 = rcp_ftz(quad_sum(sm1));
        // P as packed bf16 pairs: A fragments of P V' and, for training, the saved probabilities
        uint32_t pk[NTS][2];
#pragma unroll
        for (int nt = 0; nt < NTS; ++nt) {
          pk[nt][0] = pack2_bf16(sacc[nt][0] * inv0, sacc[nt][1] * inv0);
          pk[nt][1] = pack2_bf16(sacc[nt][2] * inv1, sacc[nt][3] * inv1);
        }
        if (p.save_p) {
          uint32_t* pp0 = reinterpret_cast<uint32_t*>(p.save_p) + w0i;
          uint32_t* pp1 = reinterpret_cast<uint32_t*>(p.save_p) + w1i;
#pragma unroll
          for (int nt = 0; nt < NTS; ++nt) {
            if (nt * 8 + 2 * t < M) {
              if (ok0) pp0[4 * nt] = pk[nt][0];
              if (ok1) pp1[4 * nt] = pk[nt][1];
            }
          }
        }
        // O += P V'.  k-step ks covers keys 16ks..16ks+15; output n-tile ot, column n <-> head-dim element 8n+ot
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
          uint32_t a[4];
          a[0] = pk[2 * ks][0]; a[1] = pk[2 * ks][1];
          if (2 * ks + 1 < NTS) { a[2] = pk[2 * ks + 1][0]; a[3] = pk[2 * ks + 1][1]; } else { a[2] = 0u; a[3] = 0u; }
#pragma unroll
          for (int ot = 0; ot < 8; ++ot) {
            const uint32_t sel = (ot & 1) ? 0x7632u : 0x5410u;
            const uint32_t b0 = __byte_perm(word(vw[ks][0], ot >> 1), word(vw[ks][1], ot >> 1), sel);
            const uint32_t b1 = __byte_perm(word(vw[ks][2], ot >> 1), word(vw[ks][3], ot >> 1), sel);
            mma_bf16(oacc[ot], a, b0, b1);
          }
        }
      }  // dirs

      // epilogue: lane owns head-dim elements [16t, 16t+16) of rows r0 and r1
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const bool ok = half ? ok1 : ok0;
        const int r = half ? r1 : r0;
        uint32_t m16 = 0u;                                   // relu gate of this lane's 16 elements
        if (ok) {
          const size_t off = ((size_t)b * N + r) * D + h * HD + 16 * t;
          const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
          uint4 sw[2] = {zero4, zero4}, vv[2] = {zero4, zero4};
          if (p.s) { sw[0] = ldg128(p.s + off); sw[1] = ldg128(p.s + off + 8); }
          if (p.residual) { vv[0] = ldg128(p.v0 + off); vv[1] = ldg128(p.v0 + off + 8); }
          uint32_t ow[8];
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            float o2[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int u = 2 * w + e;
              const float o = oacc[u & 7][(u >> 3) + 2 * half];      // element 16t+u <- tile (u%8), col 2t + u/8
              if (p.s) {
                const uint32_t sword = word(sw[w >> 2], w & 3), vword = word(vv[w >> 2], w & 3);
                const float x = __uint_as_float(e ? (sword & 0xffff0000u) : (sword << 16)) + o;
                m16 |= (x > 0.f) ? (1u << u) : 0u;
                o2[e] = __uint_as_float(e ? (vword & 0xffff0000u) : (vword << 16)) + fmaxf(x, 0.f);
              } else {
                o2[e] = o;
              }
            }
            ow[w] = pack2_bf16(o2[0], o2[1]);
          }
          *reinterpret_cast<uint4*>(p.v1 + off) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
          *reinterpret_cast<uint4*>(p.v1 + off + 8) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
        }
        if (p.gate) {
          // 64-bit gate word of (row, head): lane t of the quad owns bits [16t, 16t+16)
          const uint32_t m1 = __shfl_xor_sync(0xffffffffu, m16, 1);
          const uint32_t lo16 = (t & 1) ? m1 : m16, hi16 = (t & 1) ? m16 : m1;       // elements of lanes (t&~1) and (t|1)
          const uint32_t half32 = lo16 | (hi16 << 16);
          const uint32_t other = __shfl_xor_sync(0xffffffffu, half32, 2);
          if (t == 0 && ok) reinterpret_cast<uint2*>(p.gate)[((size_t)b * N + r) * H + h] = make_uint2(half32, other);
        }
      }
    }  // row tiles
  }  // heads
}

// ==========================================================================================
// attention backward: one CTA (4 warps) per (graph, dir, head)
struct BwdParams {
  int B, N, M, D, H, dirs, NP;       // NP = N rounded up to 16
  const void* q; const void* kv; const void* dv1; const unsigned long long* gate;
  float* p_dl; void* dq; void* dkv; void* dout;
  // fast path (attn_bwd_bf16_kernel): probabilities and rz as packed bf16 [B][dirs*H][N][MPAD]; dz = dL * rz replaces P in place
  bf16* p16; const bf16* rz16; int MPAD; float* dc;
  // explicit relation: the forward's per-pair term [B][dirs][N][M]; entries <= -1e15 are masked pairs, whose affinity received no
  // gradient (tf.where in graph_att_layer.py:98) -- dL keeps flowing to the label bias, which is added after the where
  const float* pair_bias;
};
constexpr int LDX = HD + 8;          // smem row stride of the Q / dO tiles (== 8 mod 32: conflict-free float2 reads)

template <int NTS> __host__ __device__ constexpr int tile_ld() { return (NTS * 8 % 32 == 8) ? NTS * 8 : (NTS * 8 / 32) * 32 + 40; }

template <typename T, bool S3, int NTS>
__global__ void __launch_bounds__(128) attn_bwd_kernel(const BwdParams p) {
  constexpr int LDT = tile_ld<NTS>();        // stride of the dL / P tiles, == 8 (mod 32)
  constexpr bool EX = sizeof(T) == 2;        // bf16 inputs are exact tf32 values
  constexpr int MJ = (NTS + 1) / 2;          // 16-row tiles over keys
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Qs = reinterpret_cast<float*>(smem_raw);   // [NP][LDX]
  float* dOs = Qs + p.NP * LDX;                     // [NP][LDX]   gated dO
  float* Ls = dOs + p.NP * LDX;                     // [NP][LDT]   dL
  float* Ps = Ls + p.NP * LDT;                      // [NP][LDT]   P
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int N = p.N, M = p.M, D = p.D, H = p.H, NP = p.NP;
  const int h = blockIdx.x % H, d = (blockIdx.x / H) % p.dirs, b = blockIdx.x / (H * p.dirs);
  const int dh = d * H + h;
  const int ldq = p.dirs * D, ldkv = 2 * p.dirs * D;
  const T* Q = static_cast<const T*>(p.q);
  const T* KV = static_cast<const T*>(p.kv);
  const T* dV1 = static_cast<const T*>(p.dv1);

  // ---- fill Qs, dOs (gated by the saved relu mask) with coalesced loads; direction 0 also emits dout
  for (int x0 = tid; x0 < NP * (HD / 8); x0 += 128 * 4) {
    float qv[4][8], dv[4][8];
    unsigned long long bits[4];
    // all global loads of four items first (one latency), then the gating / stores
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int x = x0 + u * 128, i = x / (HD / 8), e0 = (x % (HD / 8)) * 8;
      bits[u] = 0ull;
#pragma unroll
      for (int k = 0; k < 8; ++k) { qv[u][k] = 0.f; dv[u][k] = 0.f; }
      if (x < NP * (HD / 8) && i < N) {
        load8c<T>(Q + ((size_t)b * N + i) * ldq + d * D + h * HD + e0, qv[u]);
        load8c<T>(dV1 + ((size_t)b * N + i) * D + h * HD + e0, dv[u]);
        bits[u] = __ldg(p.gate + ((size_t)b * N + i) * H + h);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int x = x0 + u * 128, i = x / (HD / 8), e0 = (x % (HD / 8)) * 8;
      if (x >= NP * (HD / 8)) continue;
#pragma unroll
      for (int k = 0; k < 8; ++k) dv[u][k] = ((bits[u] >> (e0 + k)) & 1ull) ? dv[u][k] : 0.f;
      if (d == 0 && i < N) store_n<T>(static_cast<T*>(p.dout) + ((size_t)b * N + i) * D + h * HD + e0, dv[u], 8);
      *reinterpret_cast<float4*>(Qs + i * LDX + e0) = make_float4(qv[u][0], qv[u][1], qv[u][2], qv[u][3]);
      *reinterpret_cast<float4*>(Qs + i * LDX + e0 + 4) = make_float4(qv[u][4], qv[u][5], qv[u][6], qv[u][7]);
      *reinterpret_cast<float4*>(dOs + i * LDX + e0) = make_float4(dv[u][0], dv[u][1], dv[u][2], dv[u][3]);
      *reinterpret_cast<float4*>(dOs + i * LDX + e0 + 4) = make_float4(dv[u][4], dv[u][5], dv[u][6], dv[u][7]);
    }
  }
  // zero the key padding of the dL / P tiles (columns >= M are never written below)
  for (int x = tid; x < NP * LDT; x += 128) { Ls[x] = 0.f; Ps[x] = 0.f; }
  __syncthreads();

  // ---- phase A: per 16-row tile: dP = dO V'^T, dL = P o (dP - rowsum(P o dP)), dQ = dL K / 8
  float* P_g = p.p_dl + ((size_t)b * p.dirs * H + dh) * N * M;
  for (int mt = warp; mt * 16 < N; mt += 4) {
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    constexpr bool PACKED = sizeof(T) == 2 && !S3;
    float sacc[NTS][4];
#pragma unroll
    for (int nt = 0; nt < NTS; ++nt) { sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f; }
    uint4 kw[NTS][2];        // K rows (8nt+2t, 8nt+2t+1) at head-dim elements [8g, 8g+8): B operand of dQ = dL K
    if constexpr (PACKED) {
      // every global operand of this row tile is requested up front, packed (one memory round trip)
      const int r0c = min(r0, N - 1), r1c = min(r1, N - 1);
      uint4 dw[2][2], vw[NTS][2];
      const bf16* dp0 = reinterpret_cast<const bf16*>(dV1) + ((size_t)b * N + r0c) * D + h * HD + 8 * t;
      const bf16* dp1 = reinterpret_cast<const bf16*>(dV1) + ((size_t)b * N + r1c) * D + h * HD + 8 * t;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        dw[0][kk] = __ldg(reinterpret_cast<const uint4*>(dp0 + 32 * kk));
        dw[1][kk] = __ldg(reinterpret_cast<const uint4*>(dp1 + 32 * kk));
      }
      const unsigned long long g0 = r0 < N ? __ldg(p.gate + ((size_t)b * N + r0c) * H + h) : 0ull;
      const unsigned long long g1 = r1 < N ? __ldg(p.gate + ((size_t)b * N + r1c) * H + h) : 0ull;
#pragma unroll
      for (int nt = 0; nt < NTS; ++nt) {
        if (nt * 8 < M) {
          const bf16* vp = reinterpret_cast<const bf16*>(KV) + ((size_t)b * M + min(nt * 8 + g, M - 1)) * ldkv + (p.dirs + d) * D + h * HD + 8 * t;
          vw[nt][0] = __ldg(reinterpret_cast<const uint4*>(vp));
          vw[nt][1] = __ldg(reinterpret_cast<const uint4*>(vp + 32));
          const int j0 = min(nt * 8 + 2 * t, M - 1), j1 = min(nt * 8 + 2 * t + 1, M - 1);
          const bf16* kp = reinterpret_cast<const bf16*>(KV) + (size_t)b * M * ldkv + d * D + h * HD + 8 * g;
          kw[nt][0] = __ldg(reinterpret_cast<const uint4*>(kp + (size_t)j0 * ldkv));
          kw[nt][1] = __ldg(reinterpret_cast<const uint4*>(kp + (size_t)j1 * ldkv));
        }
      }
      // relu gate applied to the packed words: element e <-> bit e of the (row, head) gate word
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        uint32_t* w0 = reinterpret_cast<uint32_t*>(&dw[0][kk]);
        uint32_t* w1 = reinterpret_cast<uint32_t*>(&dw[1][kk]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int e = 32 * kk + 8 * t + 2 * i;
          w0[i] &= (((g0 >> e) & 1ull) ? 0x0000ffffu : 0u) | (((g0 >> (e + 1)) & 1ull) ? 0xffff0000u : 0u);
          w1[i] &= (((g1 >> e) & 1ull) ? 0x0000ffffu : 0u) | (((g1 >> (e + 1)) & 1ull) ? 0xffff0000u : 0u);
        }
      }
#pragma unroll
      for (int nt = 0; nt < NTS; ++nt) {
        if (nt * 8 < M) {
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint32_t a0[4] = {dw[0][kk].x, dw[0][kk].y, dw[0][kk].z, dw[0][kk].w};
            const uint32_t a1[4] = {dw[1][kk].x, dw[1][kk].y, dw[1][kk].z, dw[1][kk].w};
            const uint32_t v4[4] = {vw[nt][kk].x, vw[nt][kk].y, vw[nt][kk].z, vw[nt][kk].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t a[4] = {a0[i] << 16, a1[i] << 16, a0[i] & 0xffff0000u, a1[i] & 0xffff0000u};
              const uint32_t bb[2] = {v4[i] << 16, v4[i] & 0xffff0000u};
              mma_tf32(sacc[nt], a, bb);
            }
          }
        }
      }
    } else {
    // gated dO rows as A fragments, same per-lane element map as the V' rows below
    float da[16], db[16];
    {
      const int r0c = min(r0, N - 1), r1c = min(r1, N - 1);
      load_row16<T>(dV1 + ((size_t)b * N + r0c) * D + h * HD, t, da);
      load_row16<T>(dV1 + ((size_t)b * N + r1c) * D + h * HD, t, db);
      const unsigned long long g0 = r0 < N ? __ldg(p.gate + ((size_t)b * N + r0c) * H + h) : 0ull;
      const unsigned long long g1 = r1 < N ? __ldg(p.gate + ((size_t)b * N + r1c) * H + h) : 0ull;
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int e = row16_elem<T>(u, t);
        da[u] = ((g0 >> e) & 1ull) ? da[u] : 0.f;
        db[u] = ((g1 >> e) & 1ull) ? db[u] : 0.f;
      }
    }
#pragma unroll
    for (int nt = 0; nt < NTS; ++nt) {
      if (nt * 8 < M) {
        float vr[16];
        load_row16<T>(KV + ((size_t)b * M + min(nt * 8 + g, M - 1)) * ldkv + (p.dirs + d) * D + h * HD, t, vr);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          Opnd<S3, 4> a; Opnd<S3, 2> bb;
          a.template put<EX>(0, da[2 * ks]); a.template put<EX>(1, db[2 * ks]);
          a.template put<EX>(2, da[2 * ks + 1]); a.template put<EX>(3, db[2 * ks + 1]);
          bb.template put<EX>(0, vr[2 * ks]); bb.template put<EX>(1, vr[2 * ks + 1]);
          mma_acc<S3>(sacc[nt], a, bb);
        }
      }
    }
    }
    float pr[NTS][4];
    float dl0 = 0.f, dl1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NTS; ++nt) {
      const int c0 = nt * 8 + 2 * t;
      if ((M & 1) == 0) {
        const float2 z2 = make_float2(0.f, 0.f);
        const float2 p0 = (r0 < N && c0 < M) ? *reinterpret_cast<const float2*>(P_g + (size_t)r0 * M + c0) : z2;
        const float2 p1 = (r1 < N && c0 < M) ? *reinterpret_cast<const float2*>(P_g + (size_t)r1 * M + c0) : z2;
        pr[nt][0] = p0.x; pr[nt][1] = p0.y; pr[nt][2] = p1.x; pr[nt][3] = p1.y;
      } else {
        pr[nt][0] = (r0 < N && c0 < M) ? P_g[(size_t)r0 * M + c0] : 0.f;
        pr[nt][1] = (r0 < N && c0 + 1 < M) ? P_g[(size_t)r0 * M + c0 + 1] : 0.f;
        pr[nt][2] = (r1 < N && c0 < M) ? P_g[(size_t)r1 * M + c0] : 0.f;
        pr[nt][3] = (r1 < N && c0 + 1 < M) ? P_g[(size_t)r1 * M + c0 + 1] : 0.f;
      }
      dl0 += pr[nt][0] * sacc[nt][0] + pr[nt][1] * sacc[nt][1];
      dl1 += pr[nt][2] * sacc[nt][2] + pr[nt][3] * sacc[nt][3];
    }
    dl0 = quad_sum(dl0); dl1 = quad_sum(dl1);
    float dqacc[8][4];
#pragma unroll
    for (int ot = 0; ot < 8; ++ot) { dqacc[ot][0] = dqacc[ot][1] = dqacc[ot][2] = dqacc[ot][3] = 0.f; }
#pragma unroll
    for (int nt = 0; nt < NTS; ++nt) {
      const int c0 = nt * 8 + 2 * t;
      float dl[4];
      dl[0] = pr[nt][0] * (sacc[nt][0] - dl0); dl[1] = pr[nt][1] * (sacc[nt][1] - dl0);
      dl[2] = pr[nt][2] * (sacc[nt][2] - dl1); dl[3] = pr[nt][3] * (sacc[nt][3] - dl1);
      if ((M & 1) == 0) {
        if (r0 < N && c0 < M) *reinterpret_cast<float2*>(P_g + (size_t)r0 * M + c0) = make_float2(dl[0], dl[1]);
        if (r1 < N && c0 < M) *reinterpret_cast<float2*>(P_g + (size_t)r1 * M + c0) = make_float2(dl[2], dl[3]);
      } else {
        if (r0 < N) { if (c0 < M) P_g[(size_t)r0 * M + c0] = dl[0]; if (c0 + 1 < M) P_g[(size_t)r0 * M + c0 + 1] = dl[1]; }
        if (r1 < N) { if (c0 < M) P_g[(size_t)r1 * M + c0] = dl[2]; if (c0 + 1 < M) P_g[(size_t)r1 * M + c0 + 1] = dl[3]; }
      }
      if (p.pair_bias) {      // after the full dL went to global memory (label-bias gradient): masked pairs pass nothing to Q / K
        const float* pb = p.pair_bias + ((size_t)b * p.dirs + d) * N * M;
        if (!(r0 < N && c0 < M && __ldg(pb + (size_t)r0 * M + c0) > -1e15f)) dl[0] = 0.f;
        if (!(r0 < N && c0 + 1 < M && __ldg(pb + (size_t)r0 * M + c0 + 1) > -1e15f)) dl[1] = 0.f;
        if (!(r1 < N && c0 < M && __ldg(pb + (size_t)r1 * M + c0) > -1e15f)) dl[2] = 0.f;
        if (!(r1 < N && c0 + 1 < M && __ldg(pb + (size_t)r1 * M + c0 + 1) > -1e15f)) dl[3] = 0.f;
      }
      if (c0 < NTS * 8) {
        *reinterpret_cast<float2*>(Ls + r0 * LDT + c0) = make_float2(dl[0], dl[1]);
        *reinterpret_cast<float2*>(Ls + r1 * LDT + c0) = make_float2(dl[2], dl[3]);
        *reinterpret_cast<float2*>(Ps + r0 * LDT + c0) = make_float2(pr[nt][0], pr[nt][1]);
        *reinterpret_cast<float2*>(Ps + r1 * LDT + c0) = make_float2(pr[nt][2], pr[nt][3]);
      }
      if (nt * 8 < M) {
        if constexpr (PACKED) {
          const uint32_t a[4] = {__float_as_uint(dl[0]), __float_as_uint(dl[2]), __float_as_uint(dl[1]), __float_as_uint(dl[3])};
          const uint32_t k0w[4] = {kw[nt][0].x, kw[nt][0].y, kw[nt][0].z, kw[nt][0].w};
          const uint32_t k1w[4] = {kw[nt][1].x, kw[nt][1].y, kw[nt][1].z, kw[nt][1].w};
#pragma unroll
          for (int ot = 0; ot < 8; ++ot) {
            const uint32_t bb[2] = {(ot & 1) ? (k0w[ot >> 1] & 0xffff0000u) : (k0w[ot >> 1] << 16),
                                    (ot & 1) ? (k1w[ot >> 1] & 0xffff0000u) : (k1w[ot >> 1] << 16)};
            mma_tf32(dqacc[ot], a, bb);
          }
        } else {
          float ka[8], kb[8];
          const int j0 = min(nt * 8 + 2 * t, M - 1), j1 = min(nt * 8 + 2 * t + 1, M - 1);
          load8c<T>(KV + ((size_t)b * M + j0) * ldkv + d * D + h * HD + 8 * g, ka);
          load8c<T>(KV + ((size_t)b * M + j1) * ldkv + d * D + h * HD + 8 * g, kb);
          Opnd<S3, 4> a;
          a.set(0, dl[0]); a.set(1, dl[2]); a.set(2, dl[1]); a.set(3, dl[3]);
#pragma unroll
          for (int ot = 0; ot < 8; ++ot) {
            Opnd<S3, 2> bb;
            bb.template put<EX>(0, ka[ot]); bb.template put<EX>(1, kb[ot]);
            mma_acc<S3>(dqacc[ot], a, bb);
          }
        }
      }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = half ? r1 : r0;
      if (r < N) {
        float out[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) out[u] = 0.125f * dqacc[u & 7][(u >> 3) + 2 * half];
        store_n<T>(static_cast<T*>(p.dq) + ((size_t)b * N + r) * ldq + d * D + h * HD + 16 * t, out, 16);
      }
    }
  }
  __syncthreads();

  // ---- phase B: dK = dL^T Q / 8, dV' = P^T dO.  warp w owns head-dim elements [16w, 16w+16);
  // n-tile ot in {0,1}, column g <-> element 16w + 2g + ot;  k index = query row.
  float dk[MJ][2][4], dv[MJ][2][4];
#pragma unroll
  for (int jm = 0; jm < MJ; ++jm)
#pragma unroll
    for (int ot = 0; ot < 2; ++ot)
#pragma unroll
      for (int u = 0; u < 4; ++u) { dk[jm][ot][u] = 0.f; dv[jm][ot][u] = 0.f; }
  for (int ks = 0; ks * 8 < NP; ++ks) {
    const int i0 = ks * 8 + t, i1 = i0 + 4;
    const float2 q0 = *reinterpret_cast<const float2*>(Qs + i0 * LDX + 16 * warp + 2 * g);
    const float2 q1 = *reinterpret_cast<const float2*>(Qs + i1 * LDX + 16 * warp + 2 * g);
    const float2 o0 = *reinterpret_cast<const float2*>(dOs + i0 * LDX + 16 * warp + 2 * g);
    const float2 o1 = *reinterpret_cast<const float2*>(dOs + i1 * LDX + 16 * warp + 2 * g);
    Opnd<S3, 2> bq[2], bo[2];
    bq[0].template put<EX>(0, q0.x); bq[0].template put<EX>(1, q1.x); bq[1].template put<EX>(0, q0.y); bq[1].template put<EX>(1, q1.y);
    bo[0].template put<EX>(0, o0.x); bo[0].template put<EX>(1, o1.x); bo[1].template put<EX>(0, o0.y); bo[1].template put<EX>(1, o1.y);
#pragma unroll
    for (int jm = 0; jm < MJ; ++jm) {
      if (jm * 16 < M) {
        Opnd<S3, 4> al, ap;
        const int ja = jm * 16 + g, jb = min(ja + 8, LDT - 1);
        al.set(0, Ls[i0 * LDT + ja]); al.set(1, Ls[i0 * LDT + jb]); al.set(2, Ls[i1 * LDT + ja]); al.set(3, Ls[i1 * LDT + jb]);
        ap.set(0, Ps[i0 * LDT + ja]); ap.set(1, Ps[i0 * LDT + jb]); ap.set(2, Ps[i1 * LDT + ja]); ap.set(3, Ps[i1 * LDT + jb]);
#pragma unroll
        for (int ot = 0; ot < 2; ++ot) {
          mma_acc<S3>(dk[jm][ot], al, bq[ot]);
          mma_acc<S3>(dv[jm][ot], ap, bo[ot]);
        }
      }
    }
  }
  T* dKV = static_cast<T*>(p.dkv);
#pragma unroll
  for (int jm = 0; jm < MJ; ++jm) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int j = jm * 16 + g + 8 * half;
      if (j < M) {
        // C fragment (row, cols 2t, 2t+1) of tile ot <-> elements 16w + 4t + {ot, 2 + ot}
        T* rowp = dKV + ((size_t)b * M + j) * ldkv + h * HD + 16 * warp + 4 * t;
        store4<T>(rowp + d * D, 0.125f * dk[jm][0][2 * half], 0.125f * dk[jm][1][2 * half],
                  0.125f * dk[jm][0][2 * half + 1], 0.125f * dk[jm][1][2 * half + 1]);
        store4<T>(rowp + (p.dirs + d) * D, dv[jm][0][2 * half], dv[jm][1][2 * half], dv[jm][0][2 * half + 1],
                  dv[jm][1][2 * half + 1]);
      }
    }
  }
}

// ==========================================================================================
// attention backward, bf16 fast path (M <= 24 keys): one CTA (4 warps) per (graph, dir, head), bf16 m16n8k16 throughout.
//   phase A (warp <-> 16-query tile): dO = gate o dV1 (packed), dP = dO V'^T, dL = P o (dP - rowsum(P o dP)), dQ = dL K / 8.
//            dL replaces P in HBM (fp32, read by the geometry reduction); dL^T, P^T (bf16) and the gated dO go to shared memory.
//   phase B (warp <-> {dK, dV'} x key tile): dK = dL^T Q / 8, dV' = P^T dO with the query index as k: the A fragments are
//            32-bit reads of the transposed tiles, the B fragments are 8-element row pieces paired across two query rows by
//            one PRMT each (Q is copied to shared memory with cp.async while phase A runs, dO is written there by phase A).
// Same fragment conventions as the forward fast path; dL and P are rounded to bf16 only as MMA operands.
template <int NTS>
__global__ void __launch_bounds__(128, 4) attn_bwd_bf16_kernel(const BwdParams p) {
  constexpr int NKS = (NTS + 1) / 2;          // 16-key k-steps (dQ) = 16-key m-tiles (dK, dV')
  constexpr int LDO = HD + 8;                 // bf16 row stride of the gated dO tile (144 B: 16-byte aligned, conflict-free)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int N = p.N, M = p.M, D = p.D, H = p.H, NP = p.NP;
  const int LDT = NP + 8;                     // bf16 row stride of the transposed tiles
  bf16* dOs = reinterpret_cast<bf16*>(smem_raw);                      // [NP][LDO]   gated dO
  bf16* Qs = dOs + NP * LDO;                                          // [NP][LDO]   Q (cp.async, lands during phase A)
  bf16* LT = Qs + NP * LDO;                                           // [16 NKS][LDT]   dL^T
  bf16* PT = LT + 16 * NKS * LDT;                                     // [16 NKS][LDT]   P^T
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int h = blockIdx.x % H, d = (blockIdx.x / H) % p.dirs, b = blockIdx.x / (H * p.dirs);
  const int dh = d * H + h;
  const int ldq = p.dirs * D, ldkv = 2 * p.dirs * D;
  const bf16* Q = static_cast<const bf16*>(p.q);
  const bf16* KV = static_cast<const bf16*>(p.kv);
  const bf16* dV1 = static_cast<const bf16*>(p.dv1);
  uint32_t* P_g = reinterpret_cast<uint32_t*>(p.p16 + ((size_t)b * p.dirs * H + dh) * N * p.MPAD);        // packed pairs
  const uint32_t* RZ_g = reinterpret_cast<const uint32_t*>(p.rz16 + ((size_t)b * p.dirs * H + dh) * N * p.MPAD);
  const int MW = p.MPAD >> 1;                 // 32-bit words per row
  float dlsum = 0.f;                          // sum of dL over this lane's entries: the label constant's gradient
  const bf16* kbase = KV + (size_t)b * M * ldkv + d * D + h * HD;
  const bf16* vbase = KV + (size_t)b * M * ldkv + (p.dirs + d) * D + h * HD;

  // Q tile of this (dir, head) -> shared memory, asynchronously: phase B reads it exactly like the gated dO tile
  {
    const bf16* qsrc = Q + (size_t)b * N * ldq + d * D + h * HD;
    for (int x = tid; x < NP * (HD / 8); x += 128) {
      const int i = x >> 3, c8 = (x & 7) * 8;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(Qs + i * LDO + c8);
      const bf16* src = qsrc + (size_t)min(i, N - 1) * ldq + c8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // ---- phase A
  for (int mt = warp; mt * 16 < N; mt += 4) {
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    const int r0c = min(r0, N - 1), r1c = min(r1, N - 1);
    uint4 dw[2][2], vw[NTS][2], kw[NKS][4];
    {
      const bf16* dp0 = dV1 + ((size_t)b * N + r0c) * D + h * HD + 8 * t;
      const bf16* dp1 = dV1 + ((size_t)b * N + r1c) * D + h * HD + 8 * t;
      dw[0][0] = ldg128(dp0); dw[0][1] = ldg128(dp0 + 32);
      dw[1][0] = ldg128(dp1); dw[1][1] = ldg128(dp1 + 32);
#pragma unroll
      for (int nt = 0; nt < NTS; ++nt) {
        const bf16* vp = vbase + (size_t)min(nt * 8 + g, M - 1) * ldkv + 8 * t;
        vw[nt][0] = ldg128(vp); vw[nt][1] = ldg128(vp + 32);
      }
    }
    const unsigned long long g0 = r0 < N ? __ldg(p.gate + ((size_t)b * N + r0c) * H + h) : 0ull;
    const unsigned long long g1 = r1 < N ? __ldg(p.gate + ((size_t)b * N + r1c) * H + h) : 0ull;
    float pr[NTS][4];
#pragma unroll
    for (int nt = 0; nt < NTS; ++nt) {
      const int c0 = nt * 8 + 2 * t;
      const bool in0 = r0 < N && c0 < M, in1 = r1 < N && c0 < M;
      const uint32_t p0 = in0 ? P_g[r0 * MW + 4 * nt + t] : 0u, p1 = in1 ? P_g[r1 * MW + 4 * nt + t] : 0u;
      pr[nt][0] = __uint_as_float(p0 << 16); pr[nt][1] = __uint_as_float(p0 & 0xffff0000u);
      pr[nt][2] = __uint_as_float(p1 << 16); pr[nt][3] = __uint_as_float(p1 & 0xffff0000u);
    }
    // relu gate applied to the packed words: element e <-> bit e of the (row, head) gate word
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
      uint32_t w0[4] = {dw[0][kk].x, dw[0][kk].y, dw[0][kk].z, dw[0][kk].w};
      uint32_t w1[4] = {dw[1][kk].x, dw[1][kk].y, dw[1][kk].z, dw[1][kk].w};
      const uint32_t m0 = (uint32_t)(g0 >> (32 * kk + 8 * t)) & 0xffu, m1 = (uint32_t)(g1 >> (32 * kk + 8 * t)) & 0xffu;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        w0[i] &= ((m0 >> (2 * i)) & 1u ? 0x0000ffffu : 0u) | ((m0 >> (2 * i + 1)) & 1u ? 0xffff0000u : 0u);
        w1[i] &= ((m1 >> (2 * i)) & 1u ? 0x0000ffffu : 0u) | ((m1 >> (2 * i + 1)) & 1u ? 0xffff0000u : 0u);
      }
      dw[0][kk] = make_uint4(w0[0], w0[1], w0[2], w0[3]);
      dw[1][kk] = make_uint4(w1[0], w1[1], w1[2], w1[3]);
      *reinterpret_cast<uint4*>(dOs + r0 * LDO + 8 * t + 32 * kk) = dw[0][kk];
      *reinterpret_cast<uint4*>(dOs + r1 * LDO + 8 * t + 32 * kk) = dw[1][kk];
      if (d == 0) {   // the gated gradient also flows to s directly (v1 = v0 + relu(s + O_0 + O_1)): emitted once per head
        bf16* dout = static_cast<bf16*>(p.dout);
        if (r0 < N) *reinterpret_cast<uint4*>(dout + ((size_t)b * N + r0) * D + h * HD + 8 * t + 32 * kk) = dw[0][kk];
        if (r1 < N) *reinterpret_cast<uint4*>(dout + ((size_t)b * N + r1) * D + h * HD + 8 * t + 32 * kk) = dw[1][kk];
      }
    }
    // dP = dO V'^T
    float sacc[NTS][4];
#pragma unroll
    for (int nt = 0; nt < NTS; ++nt) { sacc[nt][0] = sacc[nt][1] = sacc[nt][2] = sacc[nt][3] = 0.f; }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int w0 = 2 * (s & 1);
      const uint32_t a[4] = {word(dw[0][s >> 1], w0), word(dw[1][s >> 1], w0), word(dw[0][s >> 1], w0 + 1), word(dw[1][s >> 1], w0 + 1)};
#pragma unroll
      for (int nt = 0; nt < NTS; ++nt) mma_bf16(sacc[nt], a, word(vw[nt][s >> 1], w0), word(vw[nt][s >> 1], w0 + 1));
    }
    // rz words (consumed when dz = dL * rz is written) and the K row pieces for dQ = dL K: requested now, arrive during the
    // softmax backward
    uint32_t zw[NTS][2];
#pragma unroll
    for (int nt = 0; nt < NTS; ++nt) {
      const bool inc = nt * 8 + 2 * t < M;
      zw[nt][0] = (inc && r0 < N) ? __ldg(RZ_g + r0 * MW + 4 * nt + t) : 0u;
      zw[nt][1] = (inc && r1 < N) ? __ldg(RZ_g + r1 * MW + 4 * nt + t) : 0u;
    }
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks) {
      const int j = 16 * ks + 2 * t;
      kw[ks][0] = ldg128(kbase + (size_t)min(j, M - 1) * ldkv + 8 * g);
      kw[ks][1] = ldg128(kbase + (size_t)min(j + 1, M - 1) * ldkv + 8 * g);
      kw[ks][2] = ldg128(kbase + (size_t)min(j + 8, M - 1) * ldkv + 8 * g);
      kw[ks][3] = ldg128(kbase + (size_t)min(j + 9, M - 1) * ldkv + 8 * g);
    }
    float dl0 = 0.f, dl1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NTS; ++nt) {
      dl0 += pr[nt][0] * sacc[nt][0] + pr[nt][1] * sacc[nt][1];
      dl1 += pr[nt][2] * sacc[nt][2] + pr[nt][3] * sacc[nt][3];
    }
    dl0 = quad_sum(dl0); dl1 = quad_sum(dl1);
#pragma unroll
    for (int nt = 0; nt < NTS; ++nt) {
      const int c0 = nt * 8 + 2 * t;
      float dl[4];
      dl[0] = pr[nt][0] * (sacc[nt][0] - dl0); dl[1] = pr[nt][1] * (sacc[nt][1] - dl0);
      dl[2] = pr[nt][2] * (sacc[nt][2] - dl1); dl[3] = pr[nt][3] * (sacc[nt][3] - dl1);
      dlsum += (dl[0] + dl[1]) + (dl[2] + dl[3]);
      // dz = dL * rz (gradient w.r.t. the pair_pos_fc pre-activation) replaces P in place, same packed layout
      if (r0 < N && c0 < M)
        P_g[r0 * MW + 4 * nt + t] = pack2_bf16(dl[0] * __uint_as_float(zw[nt][0] << 16), dl[1] * __uint_as_float(zw[nt][0] & 0xffff0000u));
      if (r1 < N && c0 < M)
        P_g[r1 * MW + 4 * nt + t] = pack2_bf16(dl[2] * __uint_as_float(zw[nt][1] << 16), dl[3] * __uint_as_float(zw[nt][1] & 0xffff0000u));
      // transposed bf16 copies for phase B: tile[key][query]
      LT[(c0) * LDT + r0] = __float2bfloat16_rn(dl[0]); LT[(c0 + 1) * LDT + r0] = __float2bfloat16_rn(dl[1]);
      LT[(c0) * LDT + r1] = __float2bfloat16_rn(dl[2]); LT[(c0 + 1) * LDT + r1] = __float2bfloat16_rn(dl[3]);
      PT[(c0) * LDT + r0] = __float2bfloat16_rn(pr[nt][0]); PT[(c0 + 1) * LDT + r0] = __float2bfloat16_rn(pr[nt][1]);
      PT[(c0) * LDT + r1] = __float2bfloat16_rn(pr[nt][2]); PT[(c0 + 1) * LDT + r1] = __float2bfloat16_rn(pr[nt][3]);
      sacc[nt][0] = dl[0]; sacc[nt][1] = dl[1]; sacc[nt][2] = dl[2]; sacc[nt][3] = dl[3];
    }
    // dQ = dL K / 8
    float dqacc[8][4];
#pragma unroll
    for (int ot = 0; ot < 8; ++ot) { dqacc[ot][0] = dqacc[ot][1] = dqacc[ot][2] = dqacc[ot][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks) {
      uint32_t a[4];
      a[0] = pack2_bf16(sacc[2 * ks][0], sacc[2 * ks][1]);
      a[1] = pack2_bf16(sacc[2 * ks][2], sacc[2 * ks][3]);
      if (2 * ks + 1 < NTS) {
        a[2] = pack2_bf16(sacc[2 * ks + 1][0], sacc[2 * ks + 1][1]);
        a[3] = pack2_bf16(sacc[2 * ks + 1][2], sacc[2 * ks + 1][3]);
      } else {
        a[2] = 0u; a[3] = 0u;
      }
#pragma unroll
      for (int ot = 0; ot < 8; ++ot) {
        const uint32_t sel = (ot & 1) ? 0x7632u : 0x5410u;
        const uint32_t b0 = __byte_perm(word(kw[ks][0], ot >> 1), word(kw[ks][1], ot >> 1), sel);
        const uint32_t b1 = __byte_perm(word(kw[ks][2], ot >> 1), word(kw[ks][3], ot >> 1), sel);
        mma_bf16(dqacc[ot], a, b0, b1);
      }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = half ? r1 : r0;
      if (r < N) {
        uint32_t ow[8];
#pragma unroll
        for (int w = 0; w < 8; ++w)     // element 16t+u <- tile (u%8), col 2t + u/8
          ow[w] = pack2_bf16(0.125f * dqacc[(2 * w) & 7][((2 * w) >> 3) + 2 * half], 0.125f * dqacc[(2 * w + 1) & 7][((2 * w + 1) >> 3) + 2 * half]);
        bf16* dst = static_cast<bf16*>(p.dq) + ((size_t)b * N + r) * ldq + d * D + h * HD + 16 * t;
        *reinterpret_cast<uint4*>(dst) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        *reinterpret_cast<uint4*>(dst + 8) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
      }
    }
  }
  if (p.dc) {      // d(label const) = sum dL  (zero in exact arithmetic: softmax shift invariance)
    dlsum = warp_sum(dlsum);
    if (lane == 0) atomicAdd(p.dc, dlsum);
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();

  // ---- phase B: warp = (kind, key tile).  kind 0: dK = dL^T Q / 8;  kind 1: dV' = P^T dO
  const int kind = warp & 1;
  const bf16* XT = kind ? PT : LT;
  const bf16* Bs = kind ? dOs : Qs;
  for (int jm = warp >> 1; jm * 16 < M; jm += 2) {
    float acc[8][4];
#pragma unroll
    for (int ot = 0; ot < 8; ++ot) { acc[ot][0] = acc[ot][1] = acc[ot][2] = acc[ot][3] = 0.f; }
    const bf16* xr0 = XT + (jm * 16 + g) * LDT + 2 * t;
    const bf16* xr1 = xr0 + 8 * LDT;
    for (int ks = 0; ks * 16 < NP; ++ks) {
      uint32_t a[4];
      a[0] = *reinterpret_cast<const uint32_t*>(xr0 + 16 * ks);
      a[1] = *reinterpret_cast<const uint32_t*>(xr1 + 16 * ks);
      a[2] = *reinterpret_cast<const uint32_t*>(xr0 + 16 * ks + 8);
      a[3] = *reinterpret_cast<const uint32_t*>(xr1 + 16 * ks + 8);
      uint4 bw[4];
      const int q0 = 16 * ks + 2 * t;
      bw[0] = *reinterpret_cast<const uint4*>(Bs + (q0) * LDO + 8 * g);
      bw[1] = *reinterpret_cast<const uint4*>(Bs + (q0 + 1) * LDO + 8 * g);
      bw[2] = *reinterpret_cast<const uint4*>(Bs + (q0 + 8) * LDO + 8 * g);
      bw[3] = *reinterpret_cast<const uint4*>(Bs + (q0 + 9) * LDO + 8 * g);
#pragma unroll
      for (int ot = 0; ot < 8; ++ot) {
        const uint32_t sel = (ot & 1) ? 0x7632u : 0x5410u;
        const uint32_t b0 = __byte_perm(word(bw[0], ot >> 1), word(bw[1], ot >> 1), sel);
        const uint32_t b1 = __byte_perm(word(bw[2], ot >> 1), word(bw[3], ot >> 1), sel);
        mma_bf16(acc[ot], a, b0, b1);
      }
    }
    const float sc = kind ? 1.f : 0.125f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int j = jm * 16 + g + 8 * half;
      if (j < M) {
        uint32_t ow[8];
#pragma unroll
        for (int w = 0; w < 8; ++w)
          ow[w] = pack2_bf16(sc * acc[(2 * w) & 7][((2 * w) >> 3) + 2 * half], sc * acc[(2 * w + 1) & 7][((2 * w + 1) >> 3) + 2 * half]);
        bf16* dst = static_cast<bf16*>(p.dkv) + ((size_t)b * M + j) * ldkv + (kind ? (p.dirs + d) : d) * D + h * HD + 16 * t;
        *reinterpret_cast<uint4*>(dst) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
        *reinterpret_cast<uint4*>(dst + 8) = make_uint4(ow[4], ow[5], ow[6], ow[7]);
      }
    }
  }
}

// ==========================================================================================
// geometry backward: dWg[d][e][h] += sum_pairs Emb[e] * dz[d,h],  dbg += sum dz,  dc += sum dL
struct GeoBwdParams {
  int B, N, M, H, dirs;
  const float* boxes; const float* pos_emb; WaveDiv wd; int fast;
  const float* dl; const float* gbias;
  float* dwg; long long dwg_stride; float* dbg; long long dbg_stride; float* dc;
  const bf16* dz16; int MPAD;      // DZ16 variant: dz = dL / z already formed by attn_bwd_bf16_kernel, packed bf16 [B][dirs*H][N][MPAD]
};
// One warp = one mma pipeline: per k-step it takes 8 consecutive (graph, pair) items.  D[feature, dh] += Emb^T[feature, pair] *
// dz[pair, dh] with the 64 features as 4 m-tiles (one per geometry term c: rows g = sin, g+8 = cos of wave number g) and dh as
// n-tiles.  Lane (g, t) evaluates sincos(term c, wave g) of pairs t and t+4 -- exactly its A fragments, nothing twice; the 32
// log-geometry terms of a k-step are computed one per lane and exchanged by shuffle.  3xTF32 keeps fp32 accuracy.
// FAST (bf16 training mode): SFU sincos / log / exp and one TF32 pass, like the forward fast path; the fp32 parity mode
// keeps accurate transcendentals and 3xTF32.
template <int DH, bool FAST, bool DZ16 = false>
__global__ void __launch_bounds__(256, 2) geo_bwd_kernel(const GeoBwdParams p) {
  constexpr int NTD = DH / 8;
  __shared__ float red[EMB * DH];
  __shared__ float redb[DH];
  __shared__ float redc;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int N = p.N, NM = p.N * p.M;
  const float LOGMIN = logf(1e-6f);
  for (int x = tid; x < EMB * DH; x += 256) red[x] = 0.f;
  if (tid < DH) redb[tid] = 0.f;
  if (tid == 0) redc = 0.f;
  __syncthreads();

  float acc[4][NTD][4];
  float bsum[NTD];
  float csum = 0.f;
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int nt = 0; nt < NTD; ++nt) { acc[c][nt][0] = acc[c][nt][1] = acc[c][nt][2] = acc[c][nt][3] = 0.f; }
#pragma unroll
  for (int nt = 0; nt < NTD; ++nt) bsum[nt] = 0.f;

  // k-step = 8 consecutive pairs of ONE graph (the last k-step of a graph is padded), so all index math is 32-bit
  const int KPG = (NM + 7) / 8;
  const int nsteps = p.B * KPG;
  const float winv = 1.0f / p.wd.d[g];
  for (int ks = blockIdx.x * 8 + warp; ks < nsteps; ks += gridDim.x * 8) {
    const int bq = ks / KPG, f0 = (ks - bq * KPG) * 8;
    // lane L evaluates geometry term L%4 of pair f0 + L/4
    float mine = 0.f;
    {
      const int fq = f0 + g;
      if (fq < NM && !p.pos_emb) {
        const int ip = fq / N, jp = fq - ip * N;
        const float* bx = p.boxes + (size_t)bq * N * 4;
        const float4 oi = box_terms(bx + ip * 4), oj = box_terms(bx + jp * 4);
        mine = FAST ? pair_log_term_fast(oi, oj, t) : pair_log_term(oi, oj, t);
      }
    }
    const int fA = f0 + t, fB = fA + 4;
    const bool vA = fA < NM, vB = fB < NM;
    const int bA = bq, bB = bq;
    // B fragments: dz = dL / z where z = exp(gbias) >= 1e-6, else 0    (d log(max(relu(z), 1e-6)) / dz)
    Opnd<!FAST, 2> bz[NTD];
    if constexpr (DZ16) {
      // dz was formed by the attention backward kernel; pair f = i*M + j sits at row i, column j of the [N][MPAD] slice
      const bf16* dzb = p.dz16 + (size_t)bq * DH * N * p.MPAD;
      const float invM = 1.0f / (float)p.M;
      const int iA = fdiv_small(min(fA, NM - 1), invM), iB = fdiv_small(min(fB, NM - 1), invM);
      const int oA = iA * p.MPAD + (fA - iA * p.M), oB = iB * p.MPAD + (fB - iB * p.M);
      float dzA[NTD], dzB[NTD];
#pragma unroll
      for (int nt = 0; nt < NTD; ++nt) {      // all loads of the k-step first
        const bf16* row = dzb + (size_t)(8 * nt + g) * N * p.MPAD;
        dzA[nt] = vA ? __bfloat162float(row[oA]) : 0.f;
        dzB[nt] = vB ? __bfloat162float(row[oB]) : 0.f;
      }
#pragma unroll
      for (int nt = 0; nt < NTD; ++nt) {
        bsum[nt] += dzA[nt] + dzB[nt];
        bz[nt].set(0, dzA[nt]); bz[nt].set(1, dzB[nt]);
      }
    } else {
      const float* dlb = p.dl + (size_t)bq * DH * NM;
      const float* gbb = p.gbias + (size_t)bq * DH * NM;
      float dlA[NTD], dlB[NTD], gbA[NTD], gbB[NTD];
#pragma unroll
      for (int nt = 0; nt < NTD; ++nt) {      // all loads of the k-step first
        const int off = (8 * nt + g) * NM;
        dlA[nt] = vA ? __ldg(dlb + off + fA) : 0.f; gbA[nt] = vA ? __ldg(gbb + off + fA) : 0.f;
        dlB[nt] = vB ? __ldg(dlb + off + fB) : 0.f; gbB[nt] = vB ? __ldg(gbb + off + fB) : 0.f;
      }
#pragma unroll
      for (int nt = 0; nt < NTD; ++nt) {
        csum += dlA[nt] + dlB[nt];
        const float dzA = (vA && gbA[nt] > LOGMIN) ? dlA[nt] * (FAST ? __expf(-gbA[nt]) : expf(-gbA[nt])) : 0.f;
        const float dzB = (vB && gbB[nt] > LOGMIN) ? dlB[nt] * (FAST ? __expf(-gbB[nt]) : expf(-gbB[nt])) : 0.f;
        bsum[nt] += dzA + dzB;
        bz[nt].set(0, dzA); bz[nt].set(1, dzB);
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float sA, cA, sB, cB;
      if (p.pos_emb) {
        const float* eA = p.pos_emb + ((size_t)bA * NM + fA) * EMB + 16 * c;
        const float* eB = p.pos_emb + ((size_t)bB * NM + fB) * EMB + 16 * c;
        sA = vA ? __ldg(eA + g) : 0.f; cA = vA ? __ldg(eA + 8 + g) : 0.f;
        sB = vB ? __ldg(eB + g) : 0.f; cB = vB ? __ldg(eB + 8 + g) : 0.f;
      } else {
        const float xA = 100.0f * __shfl_sync(0xffffffffu, mine, 4 * t + c);
        const float xB = 100.0f * __shfl_sync(0xffffffffu, mine, 4 * (t + 4) + c);
        if constexpr (FAST) {
          sincos_sfu(xA * winv, &sA, &cA);
          sincos_sfu(xB * winv, &sB, &cB);
        } else {
          sincos_cw(__fdiv_rn(xA, p.wd.d[g]), &sA, &cA);
          sincos_cw(__fdiv_rn(xB, p.wd.d[g]), &sB, &cB);
        }
        if (!vA) { sA = 0.f; cA = 0.f; }
        if (!vB) { sB = 0.f; cB = 0.f; }
      }
      Opnd<!FAST, 4> a;
      a.set(0, sA); a.set(1, cA); a.set(2, sB); a.set(3, cB);
#pragma unroll
      for (int nt = 0; nt < NTD; ++nt) mma_acc<!FAST>(acc[c][nt], a, bz[nt]);
    }
  }
  // ---- CTA reduction in shared memory, then one set of global atomics per CTA
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int nt = 0; nt < NTD; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e)
        atomicAdd(&red[(16 * c + g + 8 * (e >> 1)) * DH + 8 * nt + 2 * t + (e & 1)], acc[c][nt][e]);
#pragma unroll
  for (int nt = 0; nt < NTD; ++nt) {
    float v = bsum[nt];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (t == 0) atomicAdd(&redb[8 * nt + g], v);
  }
  csum = warp_sum(csum);
  if (lane == 0) atomicAdd(&redc, csum);
  __syncthreads();
  for (int x = tid; x < EMB * DH; x += 256) {
    const int e = x / DH, dh = x - e * DH, d = dh / p.H, h = dh - d * p.H;
    atomicAdd(p.dwg + (size_t)d * p.dwg_stride + e * p.H + h, red[x]);
  }
  if (tid < DH && p.dbg) {
    const int d = tid / p.H, h = tid - d * p.H;
    atomicAdd(p.dbg + (size_t)d * p.dbg_stride + h, redb[tid]);
  }
  if (!DZ16 && tid == 0 && p.dc) atomicAdd(p.dc, redc);
}

// ==========================================================================================
// stage 1 standalone: materialised pos_emb [B,M,N,64] (compat path for prepare_graph_variables)
__global__ void __launch_bounds__(256) pos_emb_kernel(const float* __restrict__ boxes, int B, int N, int M, WaveDiv wd,
                                                      float* __restrict__ out) {
  const long long total = (long long)B * M * N;
  for (long long x = (long long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(x % N), i = (int)((x / N) % M), b = (int)(x / ((long long)N * M));
    float emb[EMB];
    pair_embedding(box_terms(boxes + ((size_t)b * N + i) * 4), box_terms(boxes + ((size_t)b * N + j) * 4), wd, emb);
    float4* dst = reinterpret_cast<float4*>(out + x * EMB);
#pragma unroll
    for (int k = 0; k < EMB / 4; ++k) dst[k] = make_float4(emb[4 * k], emb[4 * k + 1], emb[4 * k + 2], emb[4 * k + 3]);
  }
}

template <typename K> int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) REGAT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return REGAT_OK;
}

template <typename T, bool S3>
int launch_fwd(const FwdParams& p, cudaStream_t st) {
  const int DH = p.dirs * p.H;
  const size_t smem = sizeof(float4) * MAX_ROIS + sizeof(float) * ((size_t)EMB * DH + 2 * DH + (size_t)DH * ROWS * p.MP);
  REGAT_REQUIRE(smem <= 227 * 1024, REGAT_ERR_UNSUPPORTED, "geoattn_fwd: bias tile needs %zu B of shared memory (nongt_dim too large)", smem);
  dim3 grid(ceil_div(p.N, ROWS), p.B);
  const int nts = ceil_div(p.M, 8);
#define REGAT_FWD_CASE(NTS)                                                     \
  {                                                                             \
    REGAT_TRY(set_smem(geoattn_fwd_kernel<T, S3, NTS>, smem));                  \
    geoattn_fwd_kernel<T, S3, NTS><<<grid, 256, smem, st>>>(p);                 \
  }
  if (nts <= 3) REGAT_FWD_CASE(3)
  else if (nts <= 5) REGAT_FWD_CASE(5)
  else if (nts <= 13) REGAT_FWD_CASE(13)
  else REGAT_REQUIRE(false, REGAT_ERR_UNSUPPORTED, "geoattn_fwd: min(nongt_dim, N) = %d > 104 keys unsupported", p.M);
#undef REGAT_FWD_CASE
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}

int launch_fwd_fast(const FastFwdParams& p, cudaStream_t st) {
  const int DH = p.dirs * p.H;
  const int rows = std::min(p.CR, p.N);
  const size_t smem = sizeof(float4) * MAX_ROIS + sizeof(float) * ((size_t)EMB * DH + 2 * DH + (size_t)DH * rows * p.MP);
  REGAT_REQUIRE(smem <= 227 * 1024, REGAT_ERR_UNSUPPORTED, "geoattn_fwd_fast: bias tile needs %zu B of shared memory", smem);
  dim3 grid(ceil_div(p.N, p.CR), p.B);
  if (p.dirs == 2) {
    REGAT_TRY(set_smem(geoattn_fwd_fast_kernel<3, 2>, smem));
    geoattn_fwd_fast_kernel<3, 2><<<grid, 256, smem, st>>>(p);
  } else {
    REGAT_TRY(set_smem(geoattn_fwd_fast_kernel<3, 1>, smem));
    geoattn_fwd_fast_kernel<3, 1><<<grid, 256, smem, st>>>(p);
  }
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}

template <typename T, bool S3, int NTS>
int launch_bwd_nts(const BwdParams& p, cudaStream_t st) {
  const size_t smem = sizeof(float) * (size_t)p.NP * (2 * LDX + 2 * tile_ld<NTS>());
  REGAT_REQUIRE(smem <= 227 * 1024, REGAT_ERR_UNSUPPORTED, "attn_bwd: needs %zu B of shared memory", smem);
  REGAT_TRY(set_smem(attn_bwd_kernel<T, S3, NTS>, smem));
  attn_bwd_kernel<T, S3, NTS><<<p.B * p.dirs * p.H, 128, smem, st>>>(p);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
int launch_bwd_fast(const BwdParams& p, cudaStream_t st) {
  constexpr int NTS = 3, NKS = 2;
  const size_t smem = sizeof(bf16) * (2 * (size_t)p.NP * (HD + 8) + 2 * (size_t)16 * NKS * (p.NP + 8));
  REGAT_TRY(set_smem(attn_bwd_bf16_kernel<NTS>, smem));
  attn_bwd_bf16_kernel<NTS><<<p.B * p.dirs * p.H, 128, smem, st>>>(p);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
template <typename T, bool S3>
int launch_bwd(const BwdParams& p, cudaStream_t st) {
  const int nts = ceil_div(p.M, 8);
  if (nts <= 3) return launch_bwd_nts<T, S3, 3>(p, st);
  if (nts <= 5) return launch_bwd_nts<T, S3, 5>(p, st);
  if (nts <= 13) return launch_bwd_nts<T, S3, 13>(p, st);
  REGAT_REQUIRE(false, REGAT_ERR_UNSUPPORTED, "attn_bwd: min(nongt_dim, N) = %d > 104 keys unsupported", p.M);
}

int check_common(int B, int N, int D, int H, int dirs, int E) {
  REGAT_REQUIRE(B > 0 && N > 0, REGAT_ERR_SHAPE, "geoattn: empty batch (B=%d, N=%d)", B, N);
  REGAT_REQUIRE(N <= MAX_ROIS, REGAT_ERR_UNSUPPORTED, "geoattn: N=%d > %d objects unsupported", N, MAX_ROIS);
  REGAT_REQUIRE(H > 0 && D == H * HD, REGAT_ERR_UNSUPPORTED, "geoattn: head dim must be 64 (rel_dim=%d, heads=%d)", D, H);
  REGAT_REQUIRE(dirs >= 1 && dirs <= 2, REGAT_ERR_SHAPE, "geoattn: dir_num must be 1 or 2 (graph_att_net.py:18)");
  REGAT_REQUIRE(E == EMB, REGAT_ERR_UNSUPPORTED, "geoattn: pos_emb_dim must be 64");
  REGAT_REQUIRE((dirs * H) % 8 == 0 && dirs * H <= MAX_DH, REGAT_ERR_UNSUPPORTED,
                "geoattn: dir_num*num_heads must be a multiple of 8 and <= 32");
  return REGAT_OK;
}

}  // namespace
}  // namespace regat

using namespace regat;

extern "C" int regat_position_embedding(const float* boxes, int B, int N, int nongt_dim, int feat_dim,
                                        const float* wave_div_host, float* pos_emb, regat_stream_t stream) {
  REGAT_REQUIRE(boxes && pos_emb && wave_div_host, REGAT_ERR_ARG, "position_embedding: null pointer");
  REGAT_REQUIRE(feat_dim == EMB, REGAT_ERR_UNSUPPORTED, "position_embedding: feat_dim must be 64");
  REGAT_REQUIRE(aligned16(pos_emb), REGAT_ERR_ALIGN, "position_embedding: output not 16-byte aligned");
  if (B <= 0 || N <= 0) return REGAT_OK;
  const int M = nongt_dim < N ? nongt_dim : N;
  WaveDiv wd;
  for (int k = 0; k < 8; ++k) wd.d[k] = wave_div_host[k];
  const long long total = (long long)B * M * N;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 8);
  pos_emb_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(boxes, B, N, M, wd, pos_emb);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}

extern "C" int regat_geoattn_fwd(int dtype, int B, int N, int nongt_dim, int D, int H, int dirs, int E,
                                 const void* q, const void* kv, const float* boxes, const float* pos_emb,
                                 const float* wave_div_host, const float* wg, int64_t wg_stride,
                                 const float* alpha_g, const float* bg, int64_t bg_stride, const float* label_c,
                                 const void* s, const void* v0, int residual, void* v1, float* save_p,
                                 float* save_gbias, uint64_t* gate, regat_stream_t stream) {
  REGAT_TRY(check_common(B, N, D, H, dirs, E));
  REGAT_REQUIRE(q && kv && wg && alpha_g && v1, REGAT_ERR_ARG, "geoattn_fwd: null pointer");
  REGAT_REQUIRE(s || !residual, REGAT_ERR_ARG, "geoattn_fwd: raw-output mode (s == NULL) has no residual");
  REGAT_REQUIRE((boxes != nullptr) != (pos_emb != nullptr), REGAT_ERR_ARG,
                "geoattn_fwd: pass exactly one of boxes / pos_emb (graph_att_net.py:42-51)");
  REGAT_REQUIRE(!boxes || wave_div_host, REGAT_ERR_ARG, "geoattn_fwd: wave_div_host missing");
  REGAT_REQUIRE(!residual || v0, REGAT_ERR_ARG, "geoattn_fwd: residual needs v0");
  REGAT_REQUIRE(aligned16(q) && aligned16(kv) && (!s || aligned16(s)) && aligned16(v1) && (!v0 || aligned16(v0)) &&
                    (!pos_emb || aligned16(pos_emb)),
                REGAT_ERR_ALIGN, "geoattn_fwd: tensors must be 16-byte aligned");
  REGAT_REQUIRE(dtype == REGAT_F32 || dtype == REGAT_BF16, REGAT_ERR_DTYPE, "geoattn_fwd: bad dtype %d", dtype);
  FwdParams p;
  p.B = B; p.N = N; p.M = nongt_dim < N ? nongt_dim : N; p.D = D; p.H = H; p.dirs = dirs;
  p.MP = bias_stride(p.M); p.residual = residual;
  p.q = q; p.kv = kv; p.boxes = boxes; p.pos_emb = pos_emb;
  for (int k = 0; k < 8; ++k) p.wd.d[k] = wave_div_host ? wave_div_host[k] : 1.f;
  p.wg = wg; p.wg_stride = wg_stride; p.alpha_g = alpha_g; p.bg = bg; p.bg_stride = bg_stride; p.label_c = label_c;
  p.s = s; p.v0 = v0; p.v1 = v1; p.save_p = save_p; p.save_gb = save_gbias;
  p.gate = reinterpret_cast<unsigned long long*>(gate);
  p.pair_bias = nullptr;
  if (dtype == REGAT_F32) return launch_fwd<float, true>(p, (cudaStream_t)stream);
  return launch_fwd<bf16, false>(p, (cudaStream_t)stream);
}

extern "C" int regat_attn_bwd(int dtype, int B, int N, int nongt_dim, int D, int H, int dirs, const void* q,
                              const void* kv, const void* dv1, const uint64_t* gate, float* p_inout_dl, void* dq,
                              void* dkv, void* dout, regat_stream_t stream) {
  REGAT_TRY(check_common(B, N, D, H, dirs, EMB));
  REGAT_REQUIRE(q && kv && dv1 && gate && p_inout_dl && dq && dkv && dout, REGAT_ERR_ARG, "attn_bwd: null pointer");
  REGAT_REQUIRE(aligned16(q) && aligned16(kv) && aligned16(dv1) && aligned16(dq) && aligned16(dkv) && aligned16(dout),
                REGAT_ERR_ALIGN, "attn_bwd: tensors must be 16-byte aligned");
  REGAT_REQUIRE(dtype == REGAT_F32 || dtype == REGAT_BF16, REGAT_ERR_DTYPE, "attn_bwd: bad dtype %d", dtype);
  BwdParams p;
  p.B = B; p.N = N; p.M = nongt_dim < N ? nongt_dim : N; p.D = D; p.H = H; p.dirs = dirs;
  p.NP = (N + 15) / 16 * 16;
  p.q = q; p.kv = kv; p.dv1 = dv1; p.gate = reinterpret_cast<const unsigned long long*>(gate);
  p.p_dl = p_inout_dl; p.dq = dq; p.dkv = dkv; p.dout = dout;
  p.p16 = nullptr; p.rz16 = nullptr; p.MPAD = 0; p.dc = nullptr; p.pair_bias = nullptr;
  if (dtype == REGAT_F32) return launch_bwd<float, true>(p, (cudaStream_t)stream);
  return launch_bwd<bf16, false>(p, (cudaStream_t)stream);
}

extern "C" int regat_geo_bwd(int B, int N, int nongt_dim, int H, int dirs, int E, const float* boxes,
                             const float* pos_emb, const float* wave_div_host, const float* dl, const float* gbias,
                             float* dwg, int64_t dwg_stride, float* dbg, int64_t dbg_stride, float* dc,
                             regat_stream_t stream) {
  return regat_geo_bwd_ex(B, N, nongt_dim, H, dirs, E, boxes, pos_emb, wave_div_host, dl, gbias, dwg, dwg_stride, dbg, dbg_stride, dc,
                          0, stream);
}

extern "C" int regat_geo_bwd_ex(int B, int N, int nongt_dim, int H, int dirs, int E, const float* boxes,
                                const float* pos_emb, const float* wave_div_host, const float* dl, const float* gbias,
                                float* dwg, int64_t dwg_stride, float* dbg, int64_t dbg_stride, float* dc, int fast_math,
                                regat_stream_t stream) {
  REGAT_REQUIRE(E == EMB, REGAT_ERR_UNSUPPORTED, "geo_bwd: pos_emb_dim must be 64");
  REGAT_REQUIRE(dl && gbias && dwg, REGAT_ERR_ARG, "geo_bwd: null pointer");
  REGAT_REQUIRE((boxes != nullptr) != (pos_emb != nullptr), REGAT_ERR_ARG, "geo_bwd: pass exactly one of boxes / pos_emb");
  REGAT_REQUIRE(!boxes || wave_div_host, REGAT_ERR_ARG, "geo_bwd: wave_div_host missing");
  if (B <= 0 || N <= 0) return REGAT_OK;
  GeoBwdParams p;
  p.B = B; p.N = N; p.M = nongt_dim < N ? nongt_dim : N; p.H = H; p.dirs = dirs;
  p.boxes = boxes; p.pos_emb = pos_emb; p.fast = (fast_math && boxes) ? 1 : 0;
  for (int k = 0; k < 8; ++k) p.wd.d[k] = wave_div_host ? wave_div_host[k] : 1.f;
  p.dl = dl; p.gbias = gbias; p.dwg = dwg; p.dwg_stride = dwg_stride; p.dbg = dbg; p.dbg_stride = dbg_stride; p.dc = dc;
  p.dz16 = nullptr; p.MPAD = 0;
  const int DH = dirs * H;
  const long long ksteps = ((long long)B * p.N * p.M + 7) / 8;
  const int blocks = (int)std::max<long long>(1, std::min<long long>((ksteps + 7) / 8, (long long)num_sms() * 2));
  cudaStream_t st = (cudaStream_t)stream;
#define REGAT_GB_CASE(X) { if (p.fast) geo_bwd_kernel<X, true><<<blocks, 256, 0, st>>>(p); else geo_bwd_kernel<X, false><<<blocks, 256, 0, st>>>(p); }
  if (DH == 32) REGAT_GB_CASE(32)
  else if (DH == 24) REGAT_GB_CASE(24)
  else if (DH == 16) REGAT_GB_CASE(16)
  else if (DH == 8) REGAT_GB_CASE(8)
  else REGAT_REQUIRE(false, REGAT_ERR_UNSUPPORTED, "geo_bwd: dir_num*num_heads must be 8, 16, 24 or 32 (got %d)", DH);
#undef REGAT_GB_CASE
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// bf16 training fast path (boxes given, M = min(nongt_dim, N) <= 24): probabilities and rz = 1/z leave the forward kernel as
// packed bf16 [B][dirs*H][N][MPAD] (MPAD = M rounded up to even); the attention backward turns P into dz = dL * rz in place;
// the geometry reduction consumes dz.
extern "C" int regat_geoattn_fast_supported(int N, int nongt_dim) {
  static const int no_fast = [] { const char* e = getenv("REGAT_ATTN_GENERIC"); return e ? atoi(e) : 0; }();
  const int M = nongt_dim < N ? nongt_dim : N;
  return (!no_fast && M >= 1 && M <= 24 && N <= MAX_ROIS) ? 1 : 0;
}

extern "C" int regat_geoattn_fwd_fast(int B, int N, int nongt_dim, int D, int H, int dirs, int E, const void* q, const void* kv,
                                      const float* boxes, const float* wave_div_host, const float* wg, int64_t wg_stride,
                                      const float* alpha_g, const float* bg, int64_t bg_stride, const float* label_c, const void* s,
                                      const void* v0, int residual, void* v1, void* save_p16, void* save_rz16, uint64_t* gate,
                                      regat_stream_t stream) {
  REGAT_TRY(check_common(B, N, D, H, dirs, E));
  REGAT_REQUIRE(q && kv && wg && alpha_g && v1 && boxes && wave_div_host, REGAT_ERR_ARG, "geoattn_fwd_fast: null pointer");
  REGAT_REQUIRE(s || !residual, REGAT_ERR_ARG, "geoattn_fwd_fast: raw-output mode (s == NULL) has no residual");
  REGAT_REQUIRE(!residual || v0, REGAT_ERR_ARG, "geoattn_fwd_fast: residual needs v0");
  REGAT_REQUIRE((save_p16 != nullptr) == (save_rz16 != nullptr), REGAT_ERR_ARG, "geoattn_fwd_fast: pass both or neither of save_p16 / save_rz16");
  REGAT_REQUIRE(aligned16(q) && aligned16(kv) && (!s || aligned16(s)) && aligned16(v1) && (!v0 || aligned16(v0)) &&
                    (!save_p16 || (aligned16(save_p16) && aligned16(save_rz16))),
                REGAT_ERR_ALIGN, "geoattn_fwd_fast: tensors must be 16-byte aligned");
  REGAT_REQUIRE(regat_geoattn_fast_supported(N, nongt_dim), REGAT_ERR_UNSUPPORTED, "geoattn_fwd_fast: needs min(nongt_dim, N) <= 24");
  FastFwdParams p;
  p.B = B; p.N = N; p.M = nongt_dim < N ? nongt_dim : N; p.D = D; p.H = H; p.dirs = dirs;
  p.MP = (p.M + 1) & ~1; p.MPAD = p.MP; p.residual = residual;
  p.CR = N <= 48 ? (N + 15) / 16 * 16 : 32;
  p.q = static_cast<const bf16*>(q); p.kv = static_cast<const bf16*>(kv); p.boxes = boxes;
  for (int k = 0; k < 8; ++k) p.wd.d[k] = wave_div_host[k];
  p.wg = wg; p.wg_stride = wg_stride; p.alpha_g = alpha_g; p.bg = bg; p.bg_stride = bg_stride; p.label_c = label_c;
  p.s = static_cast<const bf16*>(s); p.v0 = static_cast<const bf16*>(v0); p.v1 = static_cast<bf16*>(v1);
  p.save_p = static_cast<bf16*>(save_p16); p.save_rz = static_cast<bf16*>(save_rz16);
  p.gate = reinterpret_cast<unsigned long long*>(gate);
  return launch_fwd_fast(p, (cudaStream_t)stream);
}

extern "C" int regat_attn_bwd_fast(int B, int N, int nongt_dim, int D, int H, int dirs, const void* q, const void* kv, const void* dv1,
                                   const uint64_t* gate, void* p16_inout_dz, const void* rz16, void* dq, void* dkv, void* dout,
                                   float* dc, regat_stream_t stream) {
  REGAT_TRY(check_common(B, N, D, H, dirs, EMB));
  REGAT_REQUIRE(q && kv && dv1 && gate && p16_inout_dz && rz16 && dq && dkv && dout, REGAT_ERR_ARG, "attn_bwd_fast: null pointer");
  REGAT_REQUIRE(aligned16(q) && aligned16(kv) && aligned16(dv1) && aligned16(dq) && aligned16(dkv) && aligned16(dout) &&
                    aligned16(p16_inout_dz) && aligned16(rz16),
                REGAT_ERR_ALIGN, "attn_bwd_fast: tensors must be 16-byte aligned");
  REGAT_REQUIRE(regat_geoattn_fast_supported(N, nongt_dim), REGAT_ERR_UNSUPPORTED, "attn_bwd_fast: needs min(nongt_dim, N) <= 24");
  BwdParams p;
  p.B = B; p.N = N; p.M = nongt_dim < N ? nongt_dim : N; p.D = D; p.H = H; p.dirs = dirs;
  p.NP = (N + 15) / 16 * 16;
  p.q = q; p.kv = kv; p.dv1 = dv1; p.gate = reinterpret_cast<const unsigned long long*>(gate);
  p.p_dl = nullptr; p.dq = dq; p.dkv = dkv; p.dout = dout;
  p.p16 = static_cast<bf16*>(p16_inout_dz); p.rz16 = static_cast<const bf16*>(rz16); p.MPAD = (p.M + 1) & ~1; p.dc = dc;
  p.pair_bias = nullptr;
  return launch_bwd_fast(p, (cudaStream_t)stream);
}

extern "C" int regat_geo_bwd_fast(int B, int N, int nongt_dim, int H, int dirs, int E, const float* boxes, const float* wave_div_host,
                                  const void* dz16, float* dwg, int64_t dwg_stride, float* dbg, int64_t dbg_stride,
                                  regat_stream_t stream) {
  REGAT_REQUIRE(E == EMB, REGAT_ERR_UNSUPPORTED, "geo_bwd_fast: pos_emb_dim must be 64");
  REGAT_REQUIRE(dz16 && dwg && boxes && wave_div_host, REGAT_ERR_ARG, "geo_bwd_fast: null pointer");
  if (B <= 0 || N <= 0) return REGAT_OK;
  GeoBwdParams p;
  p.B = B; p.N = N; p.M = nongt_dim < N ? nongt_dim : N; p.H = H; p.dirs = dirs;
  p.boxes = boxes; p.pos_emb = nullptr; p.fast = 1;
  for (int k = 0; k < 8; ++k) p.wd.d[k] = wave_div_host[k];
  p.dl = nullptr; p.gbias = nullptr; p.dwg = dwg; p.dwg_stride = dwg_stride; p.dbg = dbg; p.dbg_stride = dbg_stride; p.dc = nullptr;
  p.dz16 = static_cast<const bf16*>(dz16); p.MPAD = (p.M + 1) & ~1;
  const int DH = dirs * H;
  const long long ksteps = (long long)B * ((p.N * p.M + 7) / 8);
  const int blocks = (int)std::max<long long>(1, std::min<long long>((ksteps + 7) / 8, (long long)num_sms() * 2));
  cudaStream_t st = (cudaStream_t)stream;
  if (DH == 32) geo_bwd_kernel<32, true, true><<<blocks, 256, 0, st>>>(p);
  else if (DH == 24) geo_bwd_kernel<24, true, true><<<blocks, 256, 0, st>>>(p);
  else if (DH == 16) geo_bwd_kernel<16, true, true><<<blocks, 256, 0, st>>>(p);
  else if (DH == 8) geo_bwd_kernel<8, true, true><<<blocks, 256, 0, st>>>(p);
  else REGAT_REQUIRE(false, REGAT_ERR_UNSUPPORTED, "geo_bwd_fast: dir_num*num_heads must be 8, 16, 24 or 32 (got %d)", DH);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// explicit relation encoders (relation_encoder.py:95-143, graph_att_net.py:56-71, graph_att_layer.py:90-102)
namespace regat {
namespace {

// pair_bias[b][d][i][j] = (sum_l adj_d[b,i,j,l] > 0) ? sum_l adj_d[b,i,j,l] w[l] + bias : -9e15,  adj_1 = adj_0 transposed in (i, j)
__global__ void __launch_bounds__(256) explicit_pair_bias_kernel(const float* __restrict__ adj, const float* __restrict__ w,
                                                                 const float* __restrict__ bias, int N, int M, int L, int dirs,
                                                                 long long total, float* __restrict__ out) {
  for (long long x = (long long)blockIdx.x * 256 + threadIdx.x; x < total; x += (long long)gridDim.x * 256) {
    const int j = (int)(x % M), i = (int)((x / M) % N), d = (int)((x / ((long long)M * N)) % dirs);
    const long long b = x / ((long long)M * N * dirs);
    const float* a = adj + ((b * N + (d ? j : i)) * N + (d ? i : j)) * L;
    float cnt = 0.f, lab = bias ? __ldg(bias) : 0.f;
    for (int l = 0; l < L; ++l) { const float v = __ldg(a + l); cnt += v; lab = fmaf(v, __ldg(w + l), lab); }
    out[x] = cnt > 0.f ? lab : -9e15f;
  }
}

// dlab[b][d][i][j] = sum_h dL[b][d][h][i][j];  dw[l] += sum adj_d[b,i,j,l] dlab;  dbias += sum dlab
__global__ void __launch_bounds__(256) explicit_pair_bias_bwd_kernel(const float* __restrict__ adj, const float* __restrict__ dl, int N, int M,
                                                                     int L, int dirs, int H, long long total, float* dw, float* dbias) {
  extern __shared__ float acc[];        // [L + 1]
  for (int l = threadIdx.x; l <= L; l += 256) acc[l] = 0.f;
  __syncthreads();
  for (long long x = (long long)blockIdx.x * 256 + threadIdx.x; x < total; x += (long long)gridDim.x * 256) {
    const int j = (int)(x % M), i = (int)((x / M) % N), d = (int)((x / ((long long)M * N)) % dirs);
    const long long b = x / ((long long)M * N * dirs);
    float g = 0.f;
    for (int h = 0; h < H; ++h) g += __ldg(dl + ((((b * dirs + d) * H + h) * N + i) * (long long)M + j));
    if (g != 0.f) {
      const float* a = adj + ((b * N + (d ? j : i)) * N + (d ? i : j)) * L;
      for (int l = 0; l < L; ++l) { const float v = __ldg(a + l); if (v != 0.f) atomicAdd(acc + l, v * g); }
      atomicAdd(acc + L, g);
    }
  }
  __syncthreads();
  for (int l = threadIdx.x; l < L; l += 256) atomicAdd(dw + l, acc[l]);
  if (threadIdx.x == 0 && dbias) atomicAdd(dbias, acc[L]);
}

}  // namespace
}  // namespace regat

extern "C" int regat_explicit_pair_bias(int B, int N, int nongt_dim, int L, int dirs, const float* adj, const float* w_label,
                                        const float* b_label, float* pair_bias, regat_stream_t stream) {
  REGAT_REQUIRE(adj && w_label && pair_bias, REGAT_ERR_ARG, "explicit_pair_bias: null pointer");
  REGAT_REQUIRE(B >= 0 && N > 0 && L > 0 && dirs >= 1 && dirs <= 2 && nongt_dim > 0, REGAT_ERR_SHAPE, "explicit_pair_bias: bad shape");
  if (B == 0) return REGAT_OK;
  const int M = nongt_dim < N ? nongt_dim : N;
  const long long total = (long long)B * dirs * N * M;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 8);
  explicit_pair_bias_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(adj, w_label, b_label, N, M, L, dirs, total, pair_bias);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}

extern "C" int regat_explicit_pair_bias_bwd(int B, int N, int nongt_dim, int L, int dirs, int H, const float* adj, const float* dl,
                                            float* dw_label, float* db_label, regat_stream_t stream) {
  REGAT_REQUIRE(adj && dl && dw_label, REGAT_ERR_ARG, "explicit_pair_bias_bwd: null pointer");
  REGAT_REQUIRE(B >= 0 && N > 0 && L > 0 && L <= 1024 && dirs >= 1 && dirs <= 2 && H > 0, REGAT_ERR_SHAPE, "explicit_pair_bias_bwd: bad shape");
  if (B == 0) return REGAT_OK;
  const int M = nongt_dim < N ? nongt_dim : N;
  const long long total = (long long)B * dirs * N * M;
  const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 2);
  explicit_pair_bias_bwd_kernel<<<blocks, 256, (L + 1) * sizeof(float), (cudaStream_t)stream>>>(adj, dl, N, M, L, dirs, H, total, dw_label,
                                                                                               db_label);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}

extern "C" int regat_graphattn_explicit_fwd(int dtype, int B, int N, int nongt_dim, int D, int H, int dirs, const void* q, const void* kv,
                                            const float* pair_bias, const void* s, const void* v0, int residual, void* v1, float* save_p,
                                            uint64_t* gate, regat_stream_t stream) {
  REGAT_TRY(check_common(B, N, D, H, dirs, EMB));
  REGAT_REQUIRE(q && kv && pair_bias && v1, REGAT_ERR_ARG, "graphattn_explicit_fwd: null pointer");
  REGAT_REQUIRE(s || !residual, REGAT_ERR_ARG, "graphattn_explicit_fwd: raw-output mode (s == NULL) has no residual");
  REGAT_REQUIRE(!residual || v0, REGAT_ERR_ARG, "graphattn_explicit_fwd: residual needs v0");
  REGAT_REQUIRE(aligned16(q) && aligned16(kv) && (!s || aligned16(s)) && aligned16(v1) && (!v0 || aligned16(v0)), REGAT_ERR_ALIGN,
                "graphattn_explicit_fwd: tensors must be 16-byte aligned");
  REGAT_REQUIRE(dtype == REGAT_F32 || dtype == REGAT_BF16, REGAT_ERR_DTYPE, "graphattn_explicit_fwd: bad dtype %d", dtype);
  FwdParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.N = N; p.M = nongt_dim < N ? nongt_dim : N; p.D = D; p.H = H; p.dirs = dirs;
  p.MP = bias_stride(p.M); p.residual = residual;
  p.q = q; p.kv = kv; p.pair_bias = pair_bias;
  p.s = s; p.v0 = v0; p.v1 = v1; p.save_p = save_p; p.gate = reinterpret_cast<unsigned long long*>(gate);
  if (dtype == REGAT_F32) return launch_fwd<float, true>(p, (cudaStream_t)stream);
  return launch_fwd<bf16, false>(p, (cudaStream_t)stream);
}

extern "C" int regat_graphattn_explicit_bwd(int dtype, int B, int N, int nongt_dim, int D, int H, int dirs, const void* q, const void* kv,
                                            const void* dv1, const uint64_t* gate, const float* pair_bias, float* p_inout_dl, void* dq,
                                            void* dkv, void* dout, regat_stream_t stream) {
  REGAT_TRY(check_common(B, N, D, H, dirs, EMB));
  REGAT_REQUIRE(q && kv && dv1 && gate && pair_bias && p_inout_dl && dq && dkv && dout, REGAT_ERR_ARG, "graphattn_explicit_bwd: null pointer");
  REGAT_REQUIRE(aligned16(q) && aligned16(kv) && aligned16(dv1) && aligned16(dq) && aligned16(dkv) && aligned16(dout), REGAT_ERR_ALIGN,
                "graphattn_explicit_bwd: tensors must be 16-byte aligned");
  REGAT_REQUIRE(dtype == REGAT_F32 || dtype == REGAT_BF16, REGAT_ERR_DTYPE, "graphattn_explicit_bwd: bad dtype %d", dtype);
  BwdParams p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.N = N; p.M = nongt_dim < N ? nongt_dim : N; p.D = D; p.H = H; p.dirs = dirs;
  p.NP = (N + 15) / 16 * 16;
  p.q = q; p.kv = kv; p.dv1 = dv1; p.gate = reinterpret_cast<const unsigned long long*>(gate);
  p.p_dl = p_inout_dl; p.dq = dq; p.dkv = dkv; p.dout = dout; p.pair_bias = pair_bias;
  if (dtype == REGAT_F32) return launch_bwd<float, true>(p, (cudaStream_t)stream);
  return launch_bwd<bf16, false>(p, (cudaStream_t)stream);
}
