// Bandwidth-bound helper kernels of the hot path: weight-norm statistics, padded-object mask,
// BUTD pooling, BCE loss, bias gradients, per-tensor clip + Adamax.  All HBM-bound: coalesced
// 128-bit accesses, grid sized in multiples of the SM count, warp-shuffle reductions.
#include <algorithm>

#include <mutex>

#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace regat {
namespace {

template <typename T> struct Vec8;   // 8 elements = 16 B (bf16) or 32 B (fp32)
template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&v)[8]) {
  if constexpr (sizeof(T) == 4) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  } else {
    uint4 x = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[2 * i] = __uint_as_float(w[i] << 16); v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
  }
}
template <typename T> __device__ __forceinline__ void st8(T* p, const float (&v)[8]) {
  if constexpr (sizeof(T) == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  } else {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]); w[k] = *reinterpret_cast<uint32_t*>(&h); }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

__device__ __forceinline__ float block_sum_256(float v, float* red) {   // blockDim.x == 256
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  if (threadIdx.x < 32) {
    s = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    s = warp_sum(s);
  }
  __syncthreads();
  return s;   // valid in warp 0
}

// Programmatic dependent launch between the optimizer's small kernels of one range: the dependent grid is scheduled while the
// upstream one still runs and blocks in pdl_wait() until that grid has completed and flushed (griddepcontrol.wait); without the
// launch attribute both calls are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------- weight norm (weight_norm.py:35-41)
constexpr int WN_CHUNK = 4096;   // elements per block
__global__ void __launch_bounds__(256) wn_prepare_kernel(const float* __restrict__ params, TensorList tl, float* sumsq,
                                                         bf16* lowp, float* partials) {
  __shared__ float red[8];
  int l = 0;
  while (l + 1 < tl.n && (int)blockIdx.x >= tl.chunk_start[l + 1]) ++l;
  const long long base = (long long)(blockIdx.x - tl.chunk_start[l]) * WN_CHUNK;
  const float* v = params + tl.off[l];
  const long long n = tl.numel[l], end = min(base + (long long)WN_CHUNK, n);
  float ss = 0.f;
  if (!lowp && (end - base) == WN_CHUNK) {
    // full chunk (tensors start 256-byte aligned, chunks are 16 KB): four independent 128-bit loads per thread
    const float4* v4 = reinterpret_cast<const float4*>(v + base);
    float4 x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) x[u] = __ldg(v4 + threadIdx.x + 256 * u);
#pragma unroll
    for (int u = 0; u < 4; ++u) ss += x[u].x * x[u].x + x[u].y * x[u].y + x[u].z * x[u].z + x[u].w * x[u].w;
  } else {
    for (long long i = base + threadIdx.x; i < end; i += 256) {
      const float x = v[i];
      ss = fmaf(x, x, ss);
      if (lowp) {
        const int cols = tl.cols[l];
        const long long r = i / cols, c = i - r * cols;
        lowp[tl.off_lowp[l] + r * tl.ld_lowp[l] + c] = __float2bfloat16_rn(x);
      }
    }
  }
  ss = block_sum_256(ss, red);
  if (threadIdx.x == 0) {
    if (partials) partials[blockIdx.x] = ss;      // deterministic: summed in chunk order by wn_alpha_kernel
    else atomicAdd(sumsq + l, ss);
  }
}

// second pass of the bf16 weight preparation: lowp = bf16(alpha_l * v) = bf16(W_eff), placed at (off_lowp, ld_lowp) so that
// layers sharing an input can sit side by side in one wide matrix (Q_0|Q_1, K_0|K_1|V'_0|V'_1, q2attention|question_embed).
__global__ void __launch_bounds__(256) wn_scaled_copy_kernel(const float* __restrict__ params, TensorList tl, int chunks,
                                                             const float* __restrict__ alpha, bf16* __restrict__ lowp) {
  pdl_wait();
  // persistent: a few fat blocks per SM walk the 16 KB chunks (4634 two-iteration blocks spent their time being scheduled)
  int l = 0;
  for (int chunk = blockIdx.x; chunk < chunks; chunk += gridDim.x) {
    while (l + 1 < tl.n && chunk >= tl.chunk_start[l + 1]) ++l;
    const long long base = (long long)(chunk - tl.chunk_start[l]) * WN_CHUNK;
    const float* v = params + tl.off[l];
    const long long n = tl.numel[l];
    const float a = alpha[tl.layer[l]];
    const int cols = tl.cols[l], ld = tl.ld_lowp[l];
    bf16* dst = lowp + tl.off_lowp[l];
    if ((cols & 7) == 0 && (ld & 7) == 0) {
      const long long end = min(base + (long long)WN_CHUNK, n);
      const long long i0 = base + threadIdx.x * 8, i1 = i0 + 256 * 8;      // WN_CHUNK = 2 x 256 x 8
      float x0[8], x1[8];
      if (i0 < end) ld8<float>(v + i0, x0);
      if (i1 < end) ld8<float>(v + i1, x1);
      if (i0 < end) {
#pragma unroll
        for (int u = 0; u < 8; ++u) x0[u] *= a;
        if (cols == ld) { st8<bf16>(dst + i0, x0); }
        else { const long long r = i0 / cols, c = i0 - r * cols; st8<bf16>(dst + r * ld + c, x0); }
      }
      if (i1 < end) {
#pragma unroll
        for (int u = 0; u < 8; ++u) x1[u] *= a;
        if (cols == ld) { st8<bf16>(dst + i1, x1); }
        else { const long long r = i1 / cols, c = i1 - r * cols; st8<bf16>(dst + r * ld + c, x1); }
      }
    } else {
      for (long long i = base + threadIdx.x; i < min(base + (long long)WN_CHUNK, n); i += 256) {
        const long long r = i / cols, c = i - r * cols;
        dst[r * ld + c] = __float2bfloat16_rn(a * v[i]);
      }
    }
  }
}

// dst[dst_off[l] + i] = src[off[l] + i]: biases of side-by-side layers gathered into one contiguous vector
__global__ void gather_kernel(const float* __restrict__ src, TensorList tl, float* __restrict__ dst) {
  for (int l = blockIdx.x; l < tl.n; l += gridDim.x)
    for (long long i = threadIdx.x; i < tl.numel[l]; i += blockDim.x) dst[tl.off_lowp[l] + i] = src[tl.off[l] + i];
}

struct AlphaExtras {          // small per-forward chores folded into the one-block alpha kernel (two launches fewer per step)
  int gather_n;
  long long g_src[8], g_dst[8], g_numel[8];       // biases of side-by-side layers -> one contiguous vector
  float* g_out;
  int label_layer;                                 // index of the label FC in alpha[]; < 0: none
  long long label_v_off, label_b_off;
  float* label_c;
};
__global__ void wn_alpha_kernel(const float* __restrict__ params, TensorList tl, float* __restrict__ sumsq,
                                const float* __restrict__ partials, float* alpha, float* inv_norm, AlphaExtras ex) {
  pdl_wait();
  pdl_trigger();
  // one warp per tensor; with `partials` the per-chunk sums are added in a fixed order (bitwise reproducible, so data-parallel
  // replicas that hold identical parameters compute identical alpha)
  const int l = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (l < tl.n) {
    float ss;
    if (partials) {
      ss = 0.f;
      for (int c = tl.chunk_start[l] + lane; c < tl.chunk_start[l + 1]; c += 32) ss += partials[c];
      ss = warp_sum(ss);
      if (lane == 0) sumsq[l] = ss;
    } else {
      ss = sumsq[l];
    }
    if (lane == 0) {
      const float inv = rsqrtf(fmaxf(ss, 1e-12f));     // tf.nn.l2_normalize epsilon
      inv_norm[l] = inv;
      const float a = params[tl.g_off[l]] * inv;
      alpha[l] = a;
      if (l == ex.label_layer)      // graph_att_net.py:71 on an all-ones adjacency
        *ex.label_c = a * params[ex.label_v_off] + (ex.label_b_off >= 0 ? params[ex.label_b_off] : 0.f);
    }
  }
  for (int k = 0; k < ex.gather_n; ++k)
    for (long long i = threadIdx.x; i < ex.g_numel[k]; i += blockDim.x) ex.g_out[ex.g_dst[k] + i] = params[ex.g_src[k] + i];
}

// ---------------------------------------------------------------- casts / elementwise
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) cast_kernel(const TI* __restrict__ in, TO* __restrict__ out, long long n8) {
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n8; i += (long long)gridDim.x * 256) {
    float v[8];
    ld8<TI>(in + i * 8, v);
    st8<TO>(out + i * 8, v);
  }
}

// mask[r] = (sum_d v[r,d] != 0)  -- relation_encoder.py:20-21.  One warp per row.
template <typename T>
__global__ void __launch_bounds__(256) rowmask_kernel(const T* __restrict__ v, int rows, int D, float* __restrict__ mask) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  float s = 0.f;
  for (int c = lane * 8; c < D; c += 256) {
    float x[8];
    ld8<T>(v + (size_t)r * D + c, x);
#pragma unroll
    for (int u = 0; u < 8; ++u) s += x[u];
  }
  s = warp_sum(s);
  if (lane == 0) mask[r] = s != 0.f ? 1.f : 0.f;
}


// x[r, :] = [ v[r, :] || mask[r] * q[r / N, :] ],  mask[r] = (sum_d v[r,d] != 0)   (relation_encoder.py:13-37)
template <typename T>
__global__ void __launch_bounds__(256) concat_vq_kernel(const T* __restrict__ v, const T* __restrict__ q, int rows, int N, int D,
                                                        int Q, T* __restrict__ out, float* __restrict__ mask) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  float s = 0.f;
  T* o = out + (size_t)r * (D + Q);
  for (int c = lane * 8; c < D; c += 256) {
    float x[8];
    ld8<T>(v + (size_t)r * D + c, x);
#pragma unroll
    for (int u = 0; u < 8; ++u) s += x[u];
    st8<T>(o + c, x);
  }
  s = warp_sum(s);
  const float m = s != 0.f ? 1.f : 0.f;
  if (lane == 0 && mask) mask[r] = m;
  for (int c = lane * 8; c < Q; c += 256) {
    float x[8];
    ld8<T>(q + (size_t)(r / N) * Q + c, x);
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] *= m;
    st8<T>(o + D + c, x);
  }
}

// out[r,c] = a[r,c]*b[r,c]   (joint = visual_embed * question_embed, fusion.py:39)
template <typename T>
__global__ void __launch_bounds__(256) mul_kernel(const T* __restrict__ a, int lda, const T* __restrict__ b, int ldb,
                                                  T* __restrict__ out, int ldo, int rows, int cols) {
  const long long n = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    out[(size_t)r * ldo + c] = from_f<T>(to_f(a[(size_t)r * lda + c]) * to_f(b[(size_t)r * ldb + c]));
  }
}
// da = dz*b, db = dz*a
template <typename T>
__global__ void __launch_bounds__(256) mul_bwd_kernel(const T* __restrict__ dz, int ldz, const T* __restrict__ a, int lda,
                                                      const T* __restrict__ b, int ldb, T* __restrict__ da, int ldda,
                                                      T* __restrict__ db, int lddb, int rows, int cols) {
  const long long n = (long long)rows * cols;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const int r = (int)(i / cols), c = (int)(i - (long long)r * cols);
    const float g = to_f(dz[(size_t)r * ldz + c]);
    da[(size_t)r * ldda + c] = from_f<T>(g * to_f(b[(size_t)r * ldb + c]));
    db[(size_t)r * lddb + c] = from_f<T>(g * to_f(a[(size_t)r * lda + c]));
  }
}

// ---------------------------------------------------------------- BUTD (fusion.py:43-54, :34)
// uw[b,c] = u[b,c] * (alpha_l * vl[c]);  cb[b] = sum_c bva[c]*uw[b,c] + bl
template <typename T>
__global__ void __launch_bounds__(256) butd_prep_kernel(const T* __restrict__ u, int ldu, const float* __restrict__ vl,
                                                        const float* __restrict__ alpha_l, const float* __restrict__ bva,
                                                        const float* __restrict__ bl, T* __restrict__ uw, float* __restrict__ cb,
                                                        int Hd) {
  __shared__ float red[8];
  const int b = blockIdx.x;
  const float al = *alpha_l;
  float s = 0.f;
  for (int c = threadIdx.x; c < Hd; c += 256) {
    const float x = to_f(u[(size_t)b * ldu + c]) * (al * vl[c]);
    const T xr = from_f<T>(x);
    uw[(size_t)b * Hd + c] = xr;
    s = fmaf(bva ? bva[c] : 0.f, to_f(xr), s);
  }
  s = block_sum_256(s, red);
  if (threadIdx.x == 0) cb[b] = s + (bl ? *bl : 0.f);
}
// duw_tot = duw + dcb[b]*bva;  du = duw_tot * wl';  dwl[c] += sum_b duw_tot*u;  dbva[c] += sum_b dcb[b]*uw[b,c];  dbl += sum_b dcb
// grid (Hd/256, row slabs); dwl / dbva / dbl are accumulated with atomics (caller zeroes them).
template <typename T>
__global__ void __launch_bounds__(256) butd_prep_bwd_kernel(const T* __restrict__ duw, const float* __restrict__ dcb,
                                                            const T* __restrict__ u, int ldu, const T* __restrict__ uw,
                                                            const float* __restrict__ vl, const float* __restrict__ alpha_l,
                                                            const float* __restrict__ bva, T* __restrict__ du, int lddu,
                                                            float* dwl, float* dbva, float* dbl, int B, int Hd) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  const int rpb = (B + gridDim.y - 1) / gridDim.y;
  const int b0 = blockIdx.y * rpb, b1 = min(B, b0 + rpb);
  if (c < Hd && b0 < b1) {
    const float wl = *alpha_l * vl[c], bv = bva ? bva[c] : 0.f;
    float sw = 0.f, sb = 0.f;
#pragma unroll 4
    for (int b = b0; b < b1; ++b) {
      const float g = to_f(duw[(size_t)b * Hd + c]) + dcb[b] * bv;
      du[(size_t)b * lddu + c] = from_f<T>(g * wl);
      sw = fmaf(g, to_f(u[(size_t)b * ldu + c]), sw);
      sb = fmaf(dcb[b], to_f(uw[(size_t)b * Hd + c]), sb);
    }
    atomicAdd(dwl + c, sw);          // gradient w.r.t. the effective [Hd,1] kernel of joint_emb.linear
    if (dbva) atomicAdd(dbva + c, sb);
  }
  if (dbl && blockIdx.x == 0 && threadIdx.x == 0 && b0 < b1) {
    float s = 0.f;
    for (int b = b0; b < b1; ++b) s += dcb[b];
    atomicAdd(dbl, s);
  }
}

// One CTA per graph: logits over objects, softmax over N (padded rows included, fusion.py:54), weighted sum.
template <typename T>
__global__ void __launch_bounds__(256) butd_pool_fwd_kernel(const T* __restrict__ v1, const T* __restrict__ weff,
                                                            const float* __restrict__ cb, float* __restrict__ att,
                                                            T* __restrict__ pooled, int N, int D) {
  extern __shared__ float sm[];      // [N] logits/att
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* vb = v1 + (size_t)b * N * D;
  const T* wb = weff + (size_t)b * D;
  for (int n = warp; n < N; n += 8) {
    float s = 0.f;
    for (int c = lane * 8; c < D; c += 256) {
      float x[8], w[8];
      ld8<T>(vb + (size_t)n * D + c, x);
      ld8<T>(wb + c, w);
#pragma unroll
      for (int u = 0; u < 8; ++u) s = fmaf(x[u], w[u], s);
    }
    s = warp_sum(s);
    if (lane == 0) sm[n] = s + cb[b];
  }
  __syncthreads();
  if (warp == 0) {
    float mx = -INFINITY;
    for (int n = lane; n < N; n += 32) mx = fmaxf(mx, sm[n]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int n = lane; n < N; n += 32) { const float e = expf(sm[n] - mx); sm[n] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int n = lane; n < N; n += 32) { const float a = sm[n] * inv; sm[n] = a; att[(size_t)b * N + n] = a; }
  }
  __syncthreads();
  for (int c = threadIdx.x * 8; c < D; c += 256 * 8) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int n = 0; n < N; ++n) {
      float x[8];
      ld8<T>(vb + (size_t)n * D + c, x);
      const float a = sm[n];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = fmaf(a, x[u], acc[u]);
    }
    st8<T>(pooled + (size_t)b * D + c, acc);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) butd_pool_bwd_kernel(const T* __restrict__ v1, const T* __restrict__ weff,
                                                            const float* __restrict__ att, const T* __restrict__ dpooled,
                                                            T* __restrict__ dv1, T* __restrict__ dweff, float* __restrict__ dcb,
                                                            int N, int D) {
  extern __shared__ float sm[];      // [2N]: att, dlogit
  float* a_s = sm; float* dl_s = sm + N;
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* vb = v1 + (size_t)b * N * D;
  const T* dp = dpooled + (size_t)b * D;
  for (int n = threadIdx.x; n < N; n += 256) a_s[n] = att[(size_t)b * N + n];
  // da[n] = <dpooled, v1[n]>
  for (int n = warp; n < N; n += 8) {
    float s = 0.f;
    for (int c = lane * 8; c < D; c += 256) {
      float x[8], g[8];
      ld8<T>(vb + (size_t)n * D + c, x);
      ld8<T>(dp + c, g);
#pragma unroll
      for (int u = 0; u < 8; ++u) s = fmaf(x[u], g[u], s);
    }
    s = warp_sum(s);
    if (lane == 0) dl_s[n] = s;
  }
  __syncthreads();
  if (warp == 0) {
    float dot = 0.f;
    for (int n = lane; n < N; n += 32) dot = fmaf(a_s[n], dl_s[n], dot);
    dot = warp_sum(dot);
    float tot = 0.f;
    for (int n = lane; n < N; n += 32) { const float d = a_s[n] * (dl_s[n] - dot); dl_s[n] = d; tot += d; }
    tot = warp_sum(tot);
    if (lane == 0) dcb[b] = tot;
  }
  __syncthreads();
  for (int c = threadIdx.x * 8; c < D; c += 256 * 8) {
    float g[8], w[8], acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    ld8<T>(dp + c, g);
    ld8<T>(weff + (size_t)b * D + c, w);
    for (int n = 0; n < N; ++n) {
      float x[8], o[8];
      ld8<T>(vb + (size_t)n * D + c, x);
      const float a = a_s[n], d = dl_s[n];
#pragma unroll
      for (int u = 0; u < 8; ++u) { o[u] = fmaf(a, g[u], d * w[u]); acc[u] = fmaf(d, x[u], acc[u]); }
      st8<T>(dv1 + ((size_t)b * N + n) * D + c, o);
    }
    st8<T>(dweff + (size_t)b * D + c, acc);
  }
}

// bf16, one graph's v1 rows fit in shared memory (N D bf16 <= 200 KB): the tile is staged ONCE with cp.async -- every 16-byte piece
// of it in flight at the same time -- and both passes over it (logits, weighted sum) read shared memory.  The kernels above walk
// the rows from global memory twice with one dependent round trip per row and warp (13.6 / 16.1 us for 19 MB; these: one round trip).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__global__ void __launch_bounds__(256) butd_pool_fwd_staged_kernel(const bf16* __restrict__ v1, const bf16* __restrict__ weff,
                                                                   const float* __restrict__ cb, float* __restrict__ att,
                                                                   bf16* __restrict__ pooled, int N, int D) {
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* vs = reinterpret_cast<bf16*>(smraw);                       // [N][D]
  float* sm = reinterpret_cast<float*>(smraw + (size_t)N * D * 2);  // [N] logits / att
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* vb = v1 + (size_t)b * N * D;
  const bf16* wb = weff + (size_t)b * D;
  for (int i = threadIdx.x; i < N * D / 8; i += 256) cp_async16(vs + (size_t)i * 8, vb + (size_t)i * 8);
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  for (int n = warp; n < N; n += 8) {
    float s = 0.f;
    for (int c = lane * 8; c < D; c += 256) {
      float x[8], w[8];
      ld8<bf16>(vs + (size_t)n * D + c, x);
      ld8<bf16>(wb + c, w);
#pragma unroll
      for (int u = 0; u < 8; ++u) s = fmaf(x[u], w[u], s);
    }
    s = warp_sum(s);
    if (lane == 0) sm[n] = s + cb[b];
  }
  __syncthreads();
  if (warp == 0) {
    float mx = -INFINITY;
    for (int n = lane; n < N; n += 32) mx = fmaxf(mx, sm[n]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int n = lane; n < N; n += 32) { const float e = expf(sm[n] - mx); sm[n] = e; sum += e; }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int n = lane; n < N; n += 32) { const float a = sm[n] * inv; sm[n] = a; att[(size_t)b * N + n] = a; }
  }
  __syncthreads();
  for (int c = threadIdx.x * 8; c < D; c += 256 * 8) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int n = 0; n < N; ++n) {
      float x[8];
      ld8<bf16>(vs + (size_t)n * D + c, x);
      const float a = sm[n];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = fmaf(a, x[u], acc[u]);
    }
    st8<bf16>(pooled + (size_t)b * D + c, acc);
  }
}
__global__ void __launch_bounds__(256) butd_pool_bwd_staged_kernel(const bf16* __restrict__ v1, const bf16* __restrict__ weff,
                                                                   const float* __restrict__ att, const bf16* __restrict__ dpooled,
                                                                   bf16* __restrict__ dv1, bf16* __restrict__ dweff,
                                                                   float* __restrict__ dcb, int N, int D) {
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* vs = reinterpret_cast<bf16*>(smraw);
  float* a_s = reinterpret_cast<float*>(smraw + (size_t)N * D * 2);
  float* dl_s = a_s + N;
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bf16* vb = v1 + (size_t)b * N * D;
  const bf16* dp = dpooled + (size_t)b * D;
  for (int i = threadIdx.x; i < N * D / 8; i += 256) cp_async16(vs + (size_t)i * 8, vb + (size_t)i * 8);
  asm volatile("cp.async.commit_group;" ::: "memory");
  for (int n = threadIdx.x; n < N; n += 256) a_s[n] = att[(size_t)b * N + n];
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  // da[n] = <dpooled, v1[n]>
  for (int n = warp; n < N; n += 8) {
    float s = 0.f;
    for (int c = lane * 8; c < D; c += 256) {
      float x[8], g[8];
      ld8<bf16>(vs + (size_t)n * D + c, x);
      ld8<bf16>(dp + c, g);
#pragma unroll
      for (int u = 0; u < 8; ++u) s = fmaf(x[u], g[u], s);
    }
    s = warp_sum(s);
    if (lane == 0) dl_s[n] = s;
  }
  __syncthreads();
  if (warp == 0) {
    float dot = 0.f;
    for (int n = lane; n < N; n += 32) dot = fmaf(a_s[n], dl_s[n], dot);
    dot = warp_sum(dot);
    float tot = 0.f;
    for (int n = lane; n < N; n += 32) { const float d = a_s[n] * (dl_s[n] - dot); dl_s[n] = d; tot += d; }
    tot = warp_sum(tot);
    if (lane == 0) dcb[b] = tot;
  }
  __syncthreads();
  // column group cg = 8 columns; the block's 256 threads cover D / 8 column groups x (256 / (D/8)) row phases
  const int groups = D / 8;
  const int phases = max(1, 256 / groups);
  const int cgi = threadIdx.x % groups, ph = threadIdx.x / groups;
  if (ph < phases) {
    const int c = cgi * 8;
    float g[8], w[8], acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    ld8<bf16>(dp + c, g);
    ld8<bf16>(weff + (size_t)b * D + c, w);
    for (int n = ph; n < N; n += phases) {
      float x[8], o[8];
      ld8<bf16>(vs + (size_t)n * D + c, x);
      const float a = a_s[n], d = dl_s[n];
#pragma unroll
      for (int u = 0; u < 8; ++u) { o[u] = fmaf(a, g[u], d * w[u]); acc[u] = fmaf(d, x[u], acc[u]); }
      st8<bf16>(dv1 + ((size_t)b * N + n) * D + c, o);
    }
    if (phases == 1) {
      st8<bf16>(dweff + (size_t)b * D + c, acc);
    } else {
      // combine the row phases of a column group through shared memory (the staged tile is no longer needed)
      __syncthreads();
      float* part = reinterpret_cast<float*>(smraw);      // [phases][D]
#pragma unroll
      for (int u = 0; u < 8; ++u) part[(size_t)ph * D + c + u] = acc[u];
      __syncthreads();
      if (ph == 0) {
        for (int q = 1; q < phases; ++q)
#pragma unroll
          for (int u = 0; u < 8; ++u) acc[u] += part[(size_t)q * D + c + u];
        st8<bf16>(dweff + (size_t)b * D + c, acc);
      }
    }
  }
}

// ---------------------------------------------------------------- loss (train.py:20-26, 107-108; score :28-39)
template <typename TD>
__global__ void __launch_bounds__(256) bce_kernel(const float* __restrict__ logits, int ldl, const float* __restrict__ target,
                                                  int A, float inv_B, float gscale, float* loss, float* score,
                                                  TD* __restrict__ dlog, int ldd) {
  __shared__ float red[8];
  __shared__ float bestv[8];
  __shared__ int besti[8];
  const int b = blockIdx.x;
  const float* x = logits + (size_t)b * ldl;
  const float* z = target + (size_t)b * A;
  float ls = 0.f, bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int a = threadIdx.x; a < ldd; a += 256) {
    float g = 0.f;
    if (a < A) {
      const float xv = x[a], zv = z[a];
      ls += fmaxf(xv, 0.f) - xv * zv + log1pf(expf(-fabsf(xv)));   // tf.nn.sigmoid_cross_entropy_with_logits
      g = (1.f / (1.f + expf(-xv)) - zv) * inv_B * gscale;
      if (xv > bv) { bv = xv; bi = a; }
    }
    if (dlog) dlog[(size_t)b * ldd + a] = from_f<TD>(g);
  }
  // argmax with first-index tie-break (np.argmax)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { bestv[threadIdx.x >> 5] = bv; besti[threadIdx.x >> 5] = bi; }
  ls = block_sum_256(ls, red);      // contains __syncthreads
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (bestv[w] > bestv[0] || (bestv[w] == bestv[0] && besti[w] < besti[0])) { bestv[0] = bestv[w]; besti[0] = besti[w]; }
    atomicAdd(loss, ls * inv_B);                 // mean over B*A, times A
    if (score) atomicAdd(score, z[besti[0]]);
  }
}

// ---------------------------------------------------------------- reductions for the backward
// out[c] += sum_r x[r,c]   (bias gradients).  A block covers 32 column groups of 8 (16 B loads for bf16) x 8 row lanes and a
// slab of rows; 4 independent row loads in flight per thread; row lanes are combined in shared memory, slabs with atomics.
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, int ld, int rows, int cols, float* out) {
  __shared__ float red[8][32][8];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + cg) * 8;
  const int rpb = (rows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rpb, r1 = min(rows, r0 + rpb);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < cols) {
    int r = r0 + rl;
    for (; r + 24 < r1; r += 32) {
      float v0[8], v1[8], v2[8], v3[8];
      ld8<T>(x + (size_t)r * ld + c, v0); ld8<T>(x + (size_t)(r + 8) * ld + c, v1);
      ld8<T>(x + (size_t)(r + 16) * ld + c, v2); ld8<T>(x + (size_t)(r + 24) * ld + c, v3);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] += (v0[u] + v1[u]) + (v2[u] + v3[u]);
    }
    for (; r < r1; r += 8) {
      float v0[8];
      ld8<T>(x + (size_t)r * ld + c, v0);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] += v0[u];
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) red[rl][cg][u] = acc[u];
  __syncthreads();
  if (rl == 0 && c < cols) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += red[k][cg][u];
      if (c + u < cols) atomicAdd(out + c + u, s);
    }
  }
}
// Several column-sum problems in one launch (the bias gradients of one backward stage).  blockIdx.y walks the row slabs of all
// problems back to back; each block sums a slab of 128 rows x 256 columns with 8 independent 16-byte loads in flight per thread.
template <typename T>
__global__ void __launch_bounds__(256) colsum_multi_kernel(ColsumBatch cb) {
  __shared__ float red[8][32][8];
  int pi = 0;
  while (pi + 1 < cb.n && (int)blockIdx.y >= cb.slab_start[pi + 1]) ++pi;
  const T* x = static_cast<const T*>(cb.x[pi]);
  const int ld = cb.ld[pi], rows = cb.rows[pi], cols = cb.cols[pi];
  float* out = cb.out[pi];
  const int cg = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c = (blockIdx.x * 32 + cg) * 8;
  if (blockIdx.x * 256 >= cols) return;
  const int r0 = ((int)blockIdx.y - cb.slab_start[pi]) * COLSUM_SLAB, r1 = min(rows, r0 + COLSUM_SLAB);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c < cols) {
    int r = r0 + rl;
    for (; r + 56 < r1; r += 64) {
      float v[8][8];
#pragma unroll
      for (int u = 0; u < 8; ++u) ld8<T>(x + (size_t)(r + 8 * u) * ld + c, v[u]);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += ((v[0][k] + v[1][k]) + (v[2][k] + v[3][k])) + ((v[4][k] + v[5][k]) + (v[6][k] + v[7][k]));
    }
    for (; r < r1; r += 8) {
      float v0[8];
      ld8<T>(x + (size_t)r * ld + c, v0);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += v0[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[rl][cg][k] = acc[k];
  __syncthreads();
  if (rl == 0 && c < cols) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += red[j][cg][k];
      if (c + k < cols) atomicAdd(out + c + k, sum);
    }
  }
}

// generic fallback (unaligned / ld not a multiple of 8)
template <typename T>
__global__ void __launch_bounds__(256) colsum_scalar_kernel(const T* __restrict__ x, int ld, int rows, int cols, float* out) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= cols) return;
  const int rpb = (rows + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rpb, r1 = min(rows, r0 + rpb);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += to_f(x[(size_t)r * ld + c]);
  if (r0 < r1) atomicAdd(out + c, s);
}
// out[b,c] = sum_n w[b*N+n] * x[(b*N+n), c]   (dsq: gradient reaching q through the mask, relation_encoder.py:31)
template <typename T>
__global__ void __launch_bounds__(256) segsum_kernel(const T* __restrict__ x, const float* __restrict__ w, int N, int D,
                                                     T* __restrict__ out) {
  const int b = blockIdx.y, c = (blockIdx.x * 256 + threadIdx.x) * 8;
  if (c >= D) return;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int n = 0; n < N; ++n) {
    const float wv = w[(size_t)b * N + n];
    if (wv != 0.f) {
      float v[8];
      ld8<T>(x + ((size_t)b * N + n) * D + c, v);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = fmaf(wv, v[u], acc[u]);
    }
  }
  st8<T>(out + (size_t)b * D + c, acc);
}
// dst[b, n, :] += src[b*M + n, :] for n < M   (keys/values are the first M objects, graph_att_layer.py:43)
template <typename T>
__global__ void __launch_bounds__(256) addrows_kernel(T* __restrict__ dst, const T* __restrict__ src, int N, int M, int D,
                                                      long long n8) {
  const int d8 = D / 8;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n8; i += (long long)gridDim.x * 256) {
    const long long row = i / d8;
    const int c = (int)(i - row * d8) * 8;
    const long long b = row / M, n = row - b * M;
    float a[8], s[8];
    T* dp = dst + ((size_t)b * N + n) * D + c;
    ld8<T>(dp, a);
    ld8<T>(src + (size_t)row * D + c, s);
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] += s[u];
    st8<T>(dp, a);
  }
}

// ---------------------------------------------------------------- optimizer (train.py:112-113, weight_norm.py:41 backward)
// pass 1: per tensor: dot = sum G*v, gg = sum G^2 (v tensors);  gg = sum g^2 (bias tensors)
__global__ void __launch_bounds__(256) opt_reduce_kernel(const float* __restrict__ params, const float* __restrict__ grads,
                                                         TensorList tl, float* stats /* per-chunk partials [chunks][2] */,
                                                         float* __restrict__ tstats /* per-tensor [n][2] */,
                                                         unsigned int* __restrict__ counters) {
  pdl_trigger();                 // the update kernel may be scheduled behind this grid right away (it waits for our results)
  __shared__ float red[8];
  __shared__ bool last;
  int l = 0;
  while (l + 1 < tl.n && (int)blockIdx.x >= tl.chunk_start[l + 1]) ++l;
  const long long base = (long long)(blockIdx.x - tl.chunk_start[l]) * WN_CHUNK;
  const long long n = tl.numel[l], end = min(base + (long long)WN_CHUNK, n);
  const float* g = grads + tl.off[l];
  const float* v = params + tl.off[l];
  float dot = 0.f, gg = 0.f;
  if (end - base == WN_CHUNK) {
    // full chunk: 2 x 4 x 16-byte loads per thread in flight before the first use.  These passes run beside the backward pass's
    // persistent GEMM CTAs, which leave room for about one such block per SM: the bandwidth has to come from loads in flight
    // per thread, not from resident blocks.
    const float4* g4 = reinterpret_cast<const float4*>(g + base);
    const float4* v4 = reinterpret_cast<const float4*>(v + base);
    float4 gs[4], vs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { gs[j] = g4[threadIdx.x + 256 * j]; vs[j] = v4[threadIdx.x + 256 * j]; }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      dot = fmaf(gs[j].x, vs[j].x, dot); dot = fmaf(gs[j].y, vs[j].y, dot); dot = fmaf(gs[j].z, vs[j].z, dot); dot = fmaf(gs[j].w, vs[j].w, dot);
      gg = fmaf(gs[j].x, gs[j].x, gg); gg = fmaf(gs[j].y, gs[j].y, gg); gg = fmaf(gs[j].z, gs[j].z, gg); gg = fmaf(gs[j].w, gs[j].w, gg);
    }
  } else {
    for (long long i = base + threadIdx.x; i < end; i += 256) {
      const float gv = g[i];
      dot = fmaf(gv, v[i], dot);
      gg = fmaf(gv, gv, gg);
    }
  }
  dot = block_sum_256(dot, red);
  gg = block_sum_256(gg, red);
  if (threadIdx.x == 0) { stats[2 * blockIdx.x] = dot; stats[2 * blockIdx.x + 1] = gg; }   // per-chunk partials
  if (!counters) return;
  // The block that finishes a tensor last adds its per-chunk partials in chunk order (bitwise reproducible, so data-parallel
  // replicas holding identical gradients derive identical statistics) -- no separate one-block kernels after this one.
  const int nchunks = tl.chunk_start[l + 1] - tl.chunk_start[l];
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int done = atomicAdd(counters + l, 1u) + 1u;
    last = done == (unsigned int)nchunks;
    if (last) counters[l] = 0u;                     // self-resetting: ready for the next step
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  float d2 = 0.f, g2 = 0.f;
  const volatile float* vs = stats;
  for (int c = tl.chunk_start[l] + (int)threadIdx.x; c < tl.chunk_start[l + 1]; c += 256) { d2 += vs[2 * c]; g2 += vs[2 * c + 1]; }
  d2 = block_sum_256(d2, red);
  g2 = block_sum_256(g2, red);
  if (threadIdx.x == 0) { tstats[2 * l] = d2; tstats[2 * l + 1] = g2; }
}

// fixed-order sum of the per-chunk partials: stats[l] = (sum G*v, sum G^2) of tensor l.  One warp per tensor.
__global__ void opt_stats_kernel(TensorList tl, const float* __restrict__ partials, float* __restrict__ stats) {
  const int l = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (l >= tl.n) return;
  float dot = 0.f, gg = 0.f;
  for (int c = tl.chunk_start[l] + lane; c < tl.chunk_start[l + 1]; c += 32) { dot += partials[2 * c]; gg += partials[2 * c + 1]; }
  dot = warp_sum(dot); gg = warp_sum(gg);
  if (lane == 0) { stats[2 * l] = dot; stats[2 * l + 1] = gg; }
}

// pass 2: tensors are listed as kind 0 (v of a weight-normed layer, grads hold dL/dW_eff), 1 (bias, plain).
// v:   dv = alpha*(G - dot*inv^2 * v),  ||dv||^2 = alpha^2 * max(gg - dot^2*inv^2, 0)
// g:   dg = dot*inv   (scalar; handled by the thread that owns element 0 of the v tensor)
// then clip_by_norm per tensor and Keras Adamax.
__global__ void __launch_bounds__(256) opt_update_kernel(float* __restrict__ params, const float* __restrict__ grads,
                                                         float* __restrict__ am, float* __restrict__ au, TensorList tl,
                                                         const float* __restrict__ stats, const float* __restrict__ alpha,
                                                         const float* __restrict__ inv_norm, OptHyper hp,
                                                         float* __restrict__ vpartials) {
  pdl_wait();                    // statistics of opt_reduce_kernel (programmatic dependent launch: we may have started early)
  pdl_trigger();
  __shared__ float red[8];
  int l = 0;
  while (l + 1 < tl.n && (int)blockIdx.x >= tl.chunk_start[l + 1]) ++l;
  const long long base = (long long)(blockIdx.x - tl.chunk_start[l]) * WN_CHUNK;
  const long long n = tl.numel[l], end = min(base + (long long)WN_CHUNK, n);
  const long long off = tl.off[l];
  const float dot = stats[2 * l], gg = stats[2 * l + 1];
  float a = 1.f, proj = 0.f, nrm;
  if (tl.kind[l] == 0 && !hp.grads_are_final) {
    const int wl = tl.layer[l];
    const float inv = inv_norm[wl];
    a = alpha[wl];
    proj = dot * inv * inv;
    // ||dv||^2 = alpha^2 (||G||^2 - <G,v>^2/||v||^2): the difference is combined in double (the two terms nearly cancel when G is
    // almost parallel to v)
    nrm = fabsf(a) * (float)sqrt(fmax((double)gg - (double)dot * (double)dot * (double)inv * (double)inv, 0.0));
  } else {
    nrm = sqrtf(gg);
  }
  const float cs = hp.clip / fmaxf(nrm, hp.clip);           // tf.clip_by_norm
  const float lr_t = hp.lr_t_dev ? __ldg(hp.lr_t_dev) : hp.lr_t;
  float ss = 0.f;
  auto step1 = [&](float w, float gr, float& m, float& u) -> float {
    const float g = cs * a * (gr - proj * w);
    m = hp.beta1 * m + (1.f - hp.beta1) * g;
    u = fmaxf(hp.beta2 * u, fabsf(g));
    const float wn = w - lr_t * m / (u + hp.eps);
    ss = fmaf(wn, wn, ss);
    return wn;
  };
  if (end - base == WN_CHUNK) {
    // full chunk: all sixteen 16-byte loads of a thread are in flight before the first use (see opt_reduce_kernel)
    float4* w4 = reinterpret_cast<float4*>(params + off + base);
    const float4* g4 = reinterpret_cast<const float4*>(grads + off + base);
    float4* m4 = reinterpret_cast<float4*>(am + off + base);
    float4* u4 = reinterpret_cast<float4*>(au + off + base);
    float4 ws[4], gs[4], ms[4], us[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = threadIdx.x + 256 * j;
      ws[j] = w4[i]; gs[j] = g4[i]; ms[j] = m4[i]; us[j] = u4[i];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = threadIdx.x + 256 * j;
      ws[j].x = step1(ws[j].x, gs[j].x, ms[j].x, us[j].x); ws[j].y = step1(ws[j].y, gs[j].y, ms[j].y, us[j].y);
      ws[j].z = step1(ws[j].z, gs[j].z, ms[j].z, us[j].z); ws[j].w = step1(ws[j].w, gs[j].w, ms[j].w, us[j].w);
      m4[i] = ms[j]; u4[i] = us[j]; w4[i] = ws[j];
    }
  } else {
    for (long long i = base + threadIdx.x; i < end; i += 256) {
      float m = am[off + i], u = au[off + i];
      const float wn = step1(params[off + i], grads[off + i], m, u);
      am[off + i] = m; au[off + i] = u;
      params[off + i] = wn;
    }
  }
  // ||v_new||^2 per chunk, laid out like wn_prepare_kernel's partials: the next forward pass skips that read of the parameters
  if (vpartials && tl.kind[l] == 0) {
    ss = block_sum_256(ss, red);
    if (threadIdx.x == 0) vpartials[tl.vchunk_start[l] + (blockIdx.x - tl.chunk_start[l])] = ss;
  }
  if (tl.kind[l] == 0 && base == 0 && threadIdx.x == 0) {    // the scalar g of this layer
    const long long go = tl.g_off[l];
    float dg = hp.grads_are_final ? grads[go] : dot * inv_norm[tl.layer[l]];
    dg *= hp.clip / fmaxf(fabsf(dg), hp.clip);
    const float m = hp.beta1 * am[go] + (1.f - hp.beta1) * dg;
    const float u = fmaxf(hp.beta2 * au[go], fabsf(dg));
    am[go] = m; au[go] = u;
    params[go] -= lr_t * m / (u + hp.eps);
  }
}

// in-place conversion of effective-weight gradients to the reference's (dv, dg)  [for inspection / parity tests]
__global__ void __launch_bounds__(256) opt_finalize_kernel(const float* __restrict__ params, float* __restrict__ grads,
                                                           TensorList tl, const float* __restrict__ stats,
                                                           const float* __restrict__ alpha, const float* __restrict__ inv_norm) {
  int l = 0;
  while (l + 1 < tl.n && (int)blockIdx.x >= tl.chunk_start[l + 1]) ++l;
  if (tl.kind[l] != 0) return;
  const long long base = (long long)(blockIdx.x - tl.chunk_start[l]) * WN_CHUNK;
  const long long n = tl.numel[l], end = min(base + (long long)WN_CHUNK, n);
  const long long off = tl.off[l];
  const int wl = tl.layer[l];
  const float dot = stats[2 * l], inv = inv_norm[wl], a = alpha[wl];
  for (long long i = base + threadIdx.x; i < end; i += 256)
    grads[off + i] = a * (grads[off + i] - dot * inv * inv * params[off + i]);
  if (base == 0 && threadIdx.x == 0) grads[tl.g_off[l]] = dot * inv;
}

__global__ void hyper_kernel(Hyper* h, float lr, int step, float beta1, int mode) {
  if (mode & HYPER_SET_LR) h->lr = lr;
  if (mode & HYPER_SET_STEP) h->step = step;
  if (mode & HYPER_TICK) hyper_tick(h, beta1);
}

__global__ void label_const_kernel(const float* __restrict__ params, long long v_off, long long b_off, const float* alpha_l,
                                   float* c) {
  *c = *alpha_l * params[v_off] + (b_off >= 0 ? params[b_off] : 0.f);   // graph_att_net.py:71 on an all-ones adjacency
}
__global__ void label_grad_kernel(const float* dc, float* grads, long long v_off, long long b_off) {
  grads[v_off] = *dc;
  if (b_off >= 0) grads[b_off] = *dc;
}

// dataset.py:329-346 (pad_sequences, padding='post') on the device: packed rows -> [B, N, width] with zero rows after each
// sample's own.  One thread per 16 bytes of OUTPUT: every output byte is written exactly once, the read side touches only
// the real rows.  width4 = width / 4.
__global__ void pad_ragged_kernel(const float4* __restrict__ packed, const int* __restrict__ offsets, int N, int width4,
                                  long long total4, long long total_rows, float4* __restrict__ padded) {
  const long long row4 = (long long)N * width4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / row4);
    const long long r = i - (long long)b * row4;
    const int n = (int)(r / width4), c = (int)(r - (long long)n * width4);
    const int lo = offsets[b], hi = offsets[b + 1];
    const bool valid = lo >= 0 && hi >= lo && hi - lo <= N && (long long)hi <= total_rows;   // else: an all-zero sample
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid && n < hi - lo) v = packed[(long long)(lo + n) * width4 + c];
    padded[i] = v;
  }
}
__global__ void pad_ragged_check_kernel(const int* __restrict__ offsets, int B, int N, long long total_rows, int* bad) {
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
    const int lo = offsets[b], hi = offsets[b + 1];
    if (lo < 0 || hi < lo || hi - lo > N || (long long)hi > total_rows) atomicExch(bad, b + 1);
  }
}

inline int grid_for(long long n, int per_block = 256) {
  return (int)std::max<long long>(1, std::min<long long>((n + per_block - 1) / per_block, (long long)num_sms() * 8));
}

}  // namespace

// ---------------------------------------------------------------- host launchers (declared in kernels.h)
// launch with the programmatic-stream-serialization attribute (REGAT_OPT_PDL=0: plain launches)
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kern)(KArgs...), int grid, int block, cudaStream_t st, Args... args) {
  static const int on = [] { const char* s = getenv("REGAT_OPT_PDL"); return s ? atoi(s) : 1; }();
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1); cfg.blockDim = dim3((unsigned)block, 1, 1); cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = (on && pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

int build_tensor_list(TensorList& tl) {
  int c = 0;
  for (int l = 0; l < tl.n; ++l) {
    tl.chunk_start[l] = c;
    c += (int)((tl.numel[l] + WN_CHUNK - 1) / WN_CHUNK);
  }
  tl.chunk_start[tl.n] = c;
  return c;
}

int k_wn_prepare(const float* params, const TensorList& tl, int chunks, float* sumsq, void* lowp, cudaStream_t st, float* partials) {
  wn_prepare_kernel<<<chunks, 256, 0, st>>>(params, tl, sumsq, static_cast<bf16*>(lowp), partials);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
int k_wn_scaled_copy(const float* params, const TensorList& tl, int chunks, const float* alpha, void* lowp, cudaStream_t st) {
  REGAT_CUDA(launch_pdl(wn_scaled_copy_kernel, std::min(chunks, num_sms() * 8), 256, st, params, tl, chunks, alpha, static_cast<bf16*>(lowp)));
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
int k_gather(const float* src, const TensorList& tl, float* dst, cudaStream_t st) {
  gather_kernel<<<tl.n, 256, 0, st>>>(src, tl, dst);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
int k_wn_alpha(const float* params, const TensorList& tl, float* sumsq, float* alpha, float* inv_norm, cudaStream_t st, const float* partials,
               const TensorList* gather, float* gather_out, int label_layer, long long label_v_off, long long label_b_off, float* label_c) {
  REGAT_REQUIRE(tl.n <= 32, REGAT_ERR_SHAPE, "wn_alpha: at most 32 tensors");
  AlphaExtras ex;
  ex.gather_n = 0; ex.g_out = gather_out; ex.label_layer = label_c ? label_layer : -1;
  ex.label_v_off = label_v_off; ex.label_b_off = label_b_off; ex.label_c = label_c;
  if (gather && gather_out) {
    REGAT_REQUIRE(gather->n <= 8, REGAT_ERR_SHAPE, "wn_alpha: at most 8 gathered biases");
    ex.gather_n = gather->n;
    for (int k = 0; k < gather->n; ++k) { ex.g_src[k] = gather->off[k]; ex.g_dst[k] = gather->off_lowp[k]; ex.g_numel[k] = gather->numel[k]; }
  }
  REGAT_CUDA(launch_pdl(wn_alpha_kernel, 1, 1024, st, params, tl, sumsq, partials, alpha, inv_norm, ex));
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
int k_cast(int to_dtype, const float* in, void* out, long long n, cudaStream_t st) {
  REGAT_REQUIRE(n % 8 == 0, REGAT_ERR_SHAPE, "cast: element count must be a multiple of 8");
  if (to_dtype == REGAT_BF16) cast_kernel<float, bf16><<<grid_for(n / 8), 256, 0, st>>>(in, static_cast<bf16*>(out), n / 8);
  else cast_kernel<float, float><<<grid_for(n / 8), 256, 0, st>>>(in, static_cast<float*>(out), n / 8);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
#define DISPATCH_T(dt, ...)                          \
  if ((dt) == REGAT_BF16) { typedef bf16 T; __VA_ARGS__; } else { typedef float T; __VA_ARGS__; }

int k_rowmask(int dt, const void* v, int rows, int D, float* mask, cudaStream_t st) {
  REGAT_REQUIRE(D % 8 == 0, REGAT_ERR_SHAPE, "rowmask: D must be a multiple of 8");
  DISPATCH_T(dt, (rowmask_kernel<T><<<ceil_div(rows, 8), 256, 0, st>>>(static_cast<const T*>(v), rows, D, mask)));
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
int k_mul(int dt, const void* a, int lda, const void* b, int ldb, void* out, int ldo, int rows, int cols, cudaStream_t st) {
  DISPATCH_T(dt, (mul_kernel<T><<<grid_for((long long)rows * cols), 256, 0, st>>>(static_cast<const T*>(a), lda, static_cast<const T*>(b), ldb,
                                                                                 static_cast<T*>(out), ldo, rows, cols)));
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
int k_mul_bwd(int dt, const void* dz, int ldz, const void* a, int lda, const void* b, int ldb, void* da, int ldda, void* db,
              int lddb, int rows, int cols, cudaStream_t st) {
  DISPATCH_T(dt, (mul_bwd_kernel<T><<<grid_for((long long)rows * cols), 256, 0, st>>>(
                     static_cast<const T*>(dz), ldz, static_cast<const T*>(a), lda, static_cast<const T*>(b), ldb,
                     static_cast<T*>(da), ldda, static_cast<T*>(db), lddb, rows, cols)));
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
int k_butd_prep(int dt, const void* u, int ldu, const float* vl, const float* alpha_l, const float* bva, const float* bl,
                void* uw, float* cb, int B, int Hd, cudaStream_t st) {
  DISPATCH_T(dt, (butd_prep_kernel<T><<<B, 256, 0, st>>>(static_cast<const T*>(u), ldu, vl, alpha_l, bva, bl, static_cast<T*>(uw), cb, Hd)));
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
int k_butd_prep_bwd(int dt, const void* duw, const float* dcb, const void* u, int ldu, const void* uw, const float* vl,
                    const float* alpha_l, const float* bva, void* du, int lddu, float* dwl, float* dbva, float* dbl, int B,
                    int Hd, cudaStream_t st) {
  DISPATCH_T(dt, (butd_prep_bwd_kernel<T><<<dim3(ceil_div(Hd, 256), std::max(1, std::min(B / 8, 32))), 256, 0, st>>>(
                     static_cast<const T*>(duw), dcb, static_cast<const T*>(u), ldu, static_cast<const T*>(uw), vl, alpha_l, bva,
                     static_cast<T*>(du), lddu, dwl, dbva, dbl, B, Hd)));
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
int k_bce(int B, int A, const float* logits, int ldl, const float* target, float gscale, float* loss, float* score,
          void* dlog, int ldd, int d_dtype, cudaStream_t st) {
  if (d_dtype == REGAT_BF16) bce_kernel<bf16><<<B, 256, 0, st>>>(logits, ldl, target, A, 1.f / B, gscale, loss, score, static_cast<bf16*>(dlog), dlog ? ldd : A);
  else bce_kernel<float><<<B, 256, 0, st>>>(logits, ldl, target, A, 1.f / B, gscale, loss, score, static_cast<float*>(dlog), dlog ? ldd : A);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
int k_colsum(int dt, const void* x, int ld, int rows, int cols, float* out, cudaStream_t st) {
  const bool vec = aligned16(x) && ld % 8 == 0 && cols % 8 == 0;
  if (vec) {
    const int gx = ceil_div(cols, 256);
    const int slabs = std::max(1, std::min(ceil_div(rows, 64), std::max(1, 4 * num_sms() / gx)));
    dim3 grid(gx, slabs);
    DISPATCH_T(dt, (colsum_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T*>(x), ld, rows, cols, out)));
  } else {
    const int slabs = std::max(1, std::min(rows / 64, 64));
    dim3 grid(ceil_div(cols, 256), slabs);
    DISPATCH_T(dt, (colsum_scalar_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T*>(x), ld, rows, cols, out)));
  }
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
// all problems must be 16-byte aligned with ld and cols multiples of 8; others go through k_colsum
int k_colsum_multi(int dt, ColsumBatch& cb, cudaStream_t st) {
  if (cb.n == 0) return REGAT_OK;
  int slabs = 0, maxcols = 0;
  for (int i = 0; i < cb.n; ++i) {
    REGAT_REQUIRE(aligned16(cb.x[i]) && cb.ld[i] % 8 == 0 && cb.cols[i] % 8 == 0, REGAT_ERR_ALIGN, "colsum_multi: unaligned problem %d", i);
    cb.slab_start[i] = slabs;
    slabs += ceil_div(cb.rows[i], COLSUM_SLAB);
    maxcols = std::max(maxcols, cb.cols[i]);
  }
  cb.slab_start[cb.n] = slabs;
  dim3 grid(ceil_div(maxcols, 256), slabs);
  DISPATCH_T(dt, (colsum_multi_kernel<T><<<grid, 256, 0, st>>>(cb)));
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
int k_segsum(int dt, const void* x, const float* w, int B, int N, int D, void* out, cudaStream_t st) {
  REGAT_REQUIRE(D % 8 == 0, REGAT_ERR_SHAPE, "segsum: D must be a multiple of 8");
  dim3 grid(ceil_div(D / 8, 256), B);
  DISPATCH_T(dt, (segsum_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T*>(x), w, N, D, static_cast<T*>(out))));
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
int k_addrows(int dt, void* dst, const void* src, int B, int N, int M, int D, cudaStream_t st) {
  REGAT_REQUIRE(D % 8 == 0, REGAT_ERR_SHAPE, "addrows: D must be a multiple of 8");
  const long long n8 = (long long)B * M * (D / 8);
  DISPATCH_T(dt, (addrows_kernel<T><<<grid_for(n8), 256, 0, st>>>(static_cast<T*>(dst), static_cast<const T*>(src), N, M, D, n8)));
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
int k_opt_reduce(const float* params, const float* grads, const TensorList& tl, int chunks, float* partials, float* stats, cudaStream_t st,
                 unsigned int* counters) {
  REGAT_REQUIRE(tl.n <= MAX_TENSORS, REGAT_ERR_SHAPE, "opt_reduce: too many tensors");
  opt_reduce_kernel<<<chunks, 256, 0, st>>>(params, grads, tl, partials, stats, counters);
  REGAT_POST_LAUNCH();
  if (counters) return REGAT_OK;
  for (int l0 = 0; l0 < tl.n; l0 += 32) {   // 32 warps (tensors) per block
    TensorList part = tl;
    if (l0) {
      for (int i = 0; i + l0 < tl.n; ++i) part.chunk_start[i] = tl.chunk_start[i + l0];
      part.chunk_start[tl.n - l0] = tl.chunk_start[tl.n];
    }
    part.n = std::min(32, tl.n - l0);
    opt_stats_kernel<<<1, 1024, 0, st>>>(part, partials, stats + 2 * l0);
    REGAT_POST_LAUNCH();
  }
  return REGAT_OK;
}
int k_opt_update(float* params, const float* grads, float* m, float* u, const TensorList& tl, int chunks, const float* stats,
                 const float* alpha, const float* inv_norm, const OptHyper& hp, cudaStream_t st, float* vpartials) {
  REGAT_CUDA(launch_pdl(opt_update_kernel, chunks, 256, st, params, grads, m, u, tl, stats, alpha, inv_norm, hp, vpartials));
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
int k_hyper(Hyper* h, float lr, int step, float beta1, int mode, cudaStream_t st) {
  hyper_kernel<<<1, 1, 0, st>>>(h, lr, step, beta1, mode);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
int k_opt_finalize(const float* params, float* grads, const TensorList& tl, int chunks, const float* stats, const float* alpha,
                   const float* inv_norm, cudaStream_t st) {
  opt_finalize_kernel<<<chunks, 256, 0, st>>>(params, grads, tl, stats, alpha, inv_norm);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
int k_label_const(const float* params, long long v_off, long long b_off, const float* alpha_l, float* c, cudaStream_t st) {
  label_const_kernel<<<1, 1, 0, st>>>(params, v_off, b_off, alpha_l, c);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
int k_label_grad(const float* dc, float* grads, long long v_off, long long b_off, cudaStream_t st) {
  label_grad_kernel<<<1, 1, 0, st>>>(dc, grads, v_off, b_off);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}

// the staged pooling kernels need more than the default 48 KB of dynamic shared memory: raised once per device to the 200 KB cap
int set_pool_smem(bool fwd, size_t /*bytes*/) {
  static std::mutex mu;
  static bool done[2][64] = {};
  int dev = 0;
  REGAT_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(mu);
  if (dev >= 0 && dev < 64 && done[fwd][dev]) return REGAT_OK;
  if (fwd) REGAT_CUDA(cudaFuncSetAttribute(butd_pool_fwd_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  else REGAT_CUDA(cudaFuncSetAttribute(butd_pool_bwd_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  if (dev >= 0 && dev < 64) done[fwd][dev] = true;
  return REGAT_OK;
}

}  // namespace regat

using namespace regat;

extern "C" int regat_butd_pool_fwd(int dtype, int B, int N, int D, const void* v1, const void* weff, const float* cb,
                                   float* att, void* pooled, regat_stream_t stream) {
  REGAT_REQUIRE(v1 && weff && cb && att && pooled, REGAT_ERR_ARG, "butd_pool_fwd: null pointer");
  REGAT_REQUIRE(D % 8 == 0, REGAT_ERR_SHAPE, "butd_pool_fwd: D must be a multiple of 8");
  REGAT_REQUIRE(aligned16(v1) && aligned16(weff) && aligned16(pooled), REGAT_ERR_ALIGN, "butd_pool_fwd: unaligned tensor");
  if (B <= 0 || N <= 0) return REGAT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  {
    const size_t staged = (size_t)N * D * 2 + (size_t)N * 4 + 16;
    static const int staged_on = [] { const char* s = getenv("REGAT_POOL_STAGED"); return s ? atoi(s) : 1; }();
    if (staged_on && dtype == REGAT_BF16 && staged <= 200 * 1024 && aligned16(v1)) {
      REGAT_TRY(set_pool_smem(true, staged));
      butd_pool_fwd_staged_kernel<<<B, 256, staged, st>>>(static_cast<const bf16*>(v1), static_cast<const bf16*>(weff), cb, att,
                                                           static_cast<bf16*>(pooled), N, D);
      REGAT_POST_LAUNCH();
      return REGAT_OK;
    }
  }
  DISPATCH_T(dtype, (butd_pool_fwd_kernel<T><<<B, 256, N * sizeof(float), st>>>(static_cast<const T*>(v1), static_cast<const T*>(weff), cb, att,
                                                                               static_cast<T*>(pooled), N, D)));
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
extern "C" int regat_butd_pool_bwd(int dtype, int B, int N, int D, const void* v1, const void* weff, const float* att,
                                   const void* dpooled, void* dv1, void* dweff, float* dcb, regat_stream_t stream) {
  REGAT_REQUIRE(v1 && weff && att && dpooled && dv1 && dweff && dcb, REGAT_ERR_ARG, "butd_pool_bwd: null pointer");
  REGAT_REQUIRE(D % 8 == 0, REGAT_ERR_SHAPE, "butd_pool_bwd: D must be a multiple of 8");
  if (B <= 0 || N <= 0) return REGAT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  {
    const int groups = D / 8, phases = std::max(1, 256 / groups);
    const size_t staged = std::max((size_t)N * D * 2, (size_t)phases * D * 4) + (size_t)2 * N * 4 + 16;
    static const int staged_on = [] { const char* s = getenv("REGAT_POOL_STAGED"); return s ? atoi(s) : 1; }();
    if (staged_on && dtype == REGAT_BF16 && staged <= 200 * 1024 && aligned16(v1) && aligned16(dv1) && aligned16(dpooled) && aligned16(weff) &&
        aligned16(dweff) && groups <= 256 && (256 % groups == 0 || groups == 256)) {
      REGAT_TRY(set_pool_smem(false, staged));
      butd_pool_bwd_staged_kernel<<<B, 256, staged, st>>>(static_cast<const bf16*>(v1), static_cast<const bf16*>(weff), att,
                                                           static_cast<const bf16*>(dpooled), static_cast<bf16*>(dv1),
                                                           static_cast<bf16*>(dweff), dcb, N, D);
      REGAT_POST_LAUNCH();
      return REGAT_OK;
    }
  }
  DISPATCH_T(dtype, (butd_pool_bwd_kernel<T><<<B, 256, 2 * N * sizeof(float), st>>>(
                        static_cast<const T*>(v1), static_cast<const T*>(weff), att, static_cast<const T*>(dpooled),
                        static_cast<T*>(dv1), static_cast<T*>(dweff), dcb, N, D)));
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
extern "C" int regat_bce_fwd_bwd(int B, int A, const float* logits, int ld_logits, const float* target, float* loss,
                                 float* score, void* dlogits, int ld_d, int d_dtype, regat_stream_t stream) {
  REGAT_REQUIRE(logits && target && loss, REGAT_ERR_ARG, "bce: null pointer");
  REGAT_REQUIRE(ld_logits >= A && (!dlogits || ld_d >= A), REGAT_ERR_SHAPE, "bce: leading dimension smaller than A");
  if (B <= 0) return REGAT_OK;
  return k_bce(B, A, logits, ld_logits, target, 1.f, loss, score, dlogits, ld_d, d_dtype, (cudaStream_t)stream);
}

extern "C" int regat_concat_visual_question(int dtype, int B, int N, int D, int Q, const void* v, const void* q, void* out,
                                            float* mask, regat_stream_t stream) {
  REGAT_REQUIRE(v && q && out, REGAT_ERR_ARG, "concat_visual_question: null pointer");
  REGAT_REQUIRE(D % 8 == 0 && Q % 8 == 0, REGAT_ERR_SHAPE, "concat_visual_question: dims must be multiples of 8");
  REGAT_REQUIRE(aligned16(v) && aligned16(q) && aligned16(out), REGAT_ERR_ALIGN, "concat_visual_question: unaligned tensor");
  if (B <= 0 || N <= 0) return REGAT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  DISPATCH_T(dtype, (concat_vq_kernel<T><<<ceil_div(B * N, 8), 256, 0, st>>>(static_cast<const T*>(v), static_cast<const T*>(q), B * N, N, D, Q,
                                                                            static_cast<T*>(out), mask)));
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
extern "C" int regat_butd_prep(int dtype, int B, int Hd, const void* u, int ldu, const float* v_linear, const float* alpha_linear,
                               const float* bias_v2att, const float* bias_linear, void* uw, float* cb, regat_stream_t stream) {
  REGAT_REQUIRE(u && v_linear && alpha_linear && uw && cb, REGAT_ERR_ARG, "butd_prep: null pointer");
  if (B <= 0) return REGAT_OK;
  return k_butd_prep(dtype, u, ldu, v_linear, alpha_linear, bias_v2att, bias_linear, uw, cb, B, Hd, (cudaStream_t)stream);
}
extern "C" int regat_mul(int dtype, int rows, int cols, const void* a, int lda, const void* b, int ldb, void* out, int ldo,
                         regat_stream_t stream) {
  REGAT_REQUIRE(a && b && out, REGAT_ERR_ARG, "mul: null pointer");
  if (rows <= 0 || cols <= 0) return REGAT_OK;
  return k_mul(dtype, a, lda, b, ldb, out, ldo, rows, cols, (cudaStream_t)stream);
}

extern "C" int regat_pad_ragged(int B, int N, int width, int64_t total_rows, const float* packed, const int32_t* offsets,
                                float* padded, int32_t* bad, regat_stream_t stream) {
  REGAT_REQUIRE(B >= 0 && N >= 0 && width > 0 && total_rows >= 0, REGAT_ERR_SHAPE, "pad_ragged: negative size");
  REGAT_REQUIRE(width % 4 == 0, REGAT_ERR_SHAPE, "pad_ragged: width (%d) must be a multiple of 4 floats", width);
  if (B == 0 || N == 0) return REGAT_OK;
  REGAT_REQUIRE(offsets && padded && (packed || total_rows == 0), REGAT_ERR_ARG, "pad_ragged: null pointer");
  REGAT_REQUIRE(aligned16(packed) && aligned16(padded), REGAT_ERR_ALIGN, "pad_ragged: buffers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (bad) {
    pad_ragged_check_kernel<<<ceil_div(B, 256), 256, 0, st>>>(offsets, B, N, (long long)total_rows, bad);
    REGAT_POST_LAUNCH();
  }
  const long long total4 = (long long)B * N * (width / 4);
  pad_ragged_kernel<<<grid_for(total4), 256, 0, st>>>(reinterpret_cast<const float4*>(packed), offsets, N, width / 4, total4,
                                                     (long long)total_rows, reinterpret_cast<float4*>(padded));
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}

extern "C" int regat_cast(int from_dtype, int to_dtype, const void* in, void* out, int64_t n, regat_stream_t stream) {
  REGAT_REQUIRE(in && out, REGAT_ERR_ARG, "cast: null pointer");
  REGAT_REQUIRE(n % 8 == 0 && aligned16(in) && aligned16(out), REGAT_ERR_ALIGN, "cast: needs 16-byte aligned buffers and a multiple of 8 elements");
  if (n == 0) return REGAT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int g = grid_for(n / 8);
  if (from_dtype == REGAT_F32 && to_dtype == REGAT_BF16) cast_kernel<float, bf16><<<g, 256, 0, st>>>(static_cast<const float*>(in), static_cast<bf16*>(out), n / 8);
  else if (from_dtype == REGAT_BF16 && to_dtype == REGAT_F32) cast_kernel<bf16, float><<<g, 256, 0, st>>>(static_cast<const bf16*>(in), static_cast<float*>(out), n / 8);
  else REGAT_REQUIRE(false, REGAT_ERR_DTYPE, "cast: only fp32 <-> bf16");
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
