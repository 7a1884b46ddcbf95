// Question front-end kernels (reference model/language_model.py:10-174), fp32.
// The dense products (input / recurrent / attention projections and every transpose of them) go through regat_gemm;
// this file holds what is not a GEMM: the masked double-table embedding and its scatter-add, the GRU gate math and
// its backward, tanh, the self-attention whose softmax runs over the BATCH axis followed by a raw reshape
// (language_model.py:163-167), the weighted pooling, and per-tensor weight-norm / clip / Adamax for the front-end's
// own parameter buffer.  Sequenced by tf_vqa_regat_b200/question.py.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace regat {
namespace {

inline int grid1d(long long n, int per_block = 256) {
  return (int)std::max<long long>(1, std::min<long long>((n + per_block - 1) / per_block, (long long)num_sms() * 8));
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// language_model.py:33-40 (+ :88-90 for op 'c'): out[r, :] = [emb[tok] | emb2[tok]] * (tok != n_token)
__global__ void q_embed_fwd_kernel(const int* __restrict__ tokens, long long BT, int n_token, int E, const float* __restrict__ emb,
                                   const float* __restrict__ emb2, float* __restrict__ out) {
  const int W = emb2 ? 2 * E : E;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < BT * W; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / W;
    const int c = (int)(i - r * W);
    const int tok = tokens[r];
    float v = 0.f;
    if (tok != n_token && tok >= 0 && tok <= n_token) v = c < E ? emb[(long long)tok * E + c] : emb2[(long long)tok * E + (c - E)];
    out[i] = v;
  }
}
// transpose of the gather: d table[tok] += dX[r] for unmasked rows (tables are zeroed by the caller)
__global__ void q_embed_bwd_kernel(const int* __restrict__ tokens, long long BT, int n_token, int E, int W,
                                   const float* __restrict__ dX, float* demb, float* demb2) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < BT * W; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / W;
    const int c = (int)(i - r * W);
    const int tok = tokens[r];
    if (tok == n_token || tok < 0 || tok > n_token) continue;
    if (c < E) { if (demb) atomicAdd(demb + (long long)tok * E + c, dX[i]); }
    else if (demb2) atomicAdd(demb2 + (long long)tok * E + (c - E), dX[i]);
  }
}

// Keras GRU step, reset_after=True, gate blocks z | r | h (language_model.py:106-108):
//   z = sig(xz + hz), r = sig(xr + hr), c = tanh(xh + r * hh), h = z * hp + (1 - z) * c
__global__ void q_gru_gates_fwd_kernel(int B, int H, const float* __restrict__ xi, long long ld_xi, const float* __restrict__ hi,
                                       const float* __restrict__ hp, long long ld_hp, float* __restrict__ h_out, long long ld_h,
                                       float* __restrict__ z_s, float* __restrict__ r_s, float* __restrict__ c_s,
                                       float* __restrict__ hp_copy) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)B * H; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / H), j = (int)(i - (long long)b * H);
    const float* x = xi + b * ld_xi;
    const float* h3 = hi + (long long)b * 3 * H;
    const float hprev = hp ? hp[b * ld_hp + j] : 0.f;
    const float z = sigmoidf_(x[j] + h3[j]);
    const float r = sigmoidf_(x[H + j] + h3[H + j]);
    const float c = tanhf(x[2 * H + j] + r * h3[2 * H + j]);
    h_out[b * ld_h + j] = z * hprev + (1.f - z) * c;
    z_s[i] = z; r_s[i] = r; c_s[i] = c;
    if (hp_copy) hp_copy[i] = hprev;
  }
}
// dh = dh_seq + dh_rec.  Writes dxi = (dz_pre, dr_pre, da), dhi = (dz_pre, dr_pre, da * r), dhp = dh * z
// (the recurrent product dhi . U^T is accumulated onto dhp by the caller's GEMM).
__global__ void q_gru_gates_bwd_kernel(int B, int H, const float* __restrict__ dh_seq, long long ld_dseq, const float* __restrict__ dh_rec,
                                       const float* __restrict__ z_s, const float* __restrict__ r_s, const float* __restrict__ c_s,
                                       const float* __restrict__ hp_copy, const float* __restrict__ hi, float* __restrict__ dxi,
                                       long long ld_dxi, float* __restrict__ dhi, float* __restrict__ dhp) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)B * H; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / H), j = (int)(i - (long long)b * H);
    const float dh = dh_seq[b * ld_dseq + j] + (dh_rec ? dh_rec[i] : 0.f);
    const float z = z_s[i], r = r_s[i], c = c_s[i], hprev = hp_copy[i];
    const float hh = hi[(long long)b * 3 * H + 2 * H + j];
    const float da = dh * (1.f - z) * (1.f - c * c);
    const float dzp = dh * (hprev - c) * z * (1.f - z);
    const float drp = da * hh * r * (1.f - r);
    float* dx = dxi + b * ld_dxi;
    float* dr3 = dhi + (long long)b * 3 * H;
    dx[j] = dzp; dx[H + j] = drp; dx[2 * H + j] = da;
    dr3[j] = dzp; dr3[H + j] = drp; dr3[2 * H + j] = da * r;
    dhp[i] = dh * z;
  }
}

__global__ void q_tanh_fwd_kernel(float* x, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) x[i] = tanhf(x[i]);
}
__global__ void q_tanh_bwd_kernel(float* dy, const float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dy[i] *= 1.f - y[i] * y[i];
}

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sh) {
  v = is_max ? warp_max(v) : warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float t = l < nw ? sh[l] : (is_max ? -INFINITY : 0.f);
  t = is_max ? warp_max(t) : warp_sum(t);
  return t;   // valid in every lane of every warp (each warp reduces the same nw values)
}
// language_model.py:163-165: P = softmax(transpose(logits [B,T]) -> [T,B], axis=1): one block per position t, over the batch.
__global__ void q_batch_softmax_fwd_kernel(const float* __restrict__ logits, int B, int T, float* __restrict__ P) {
  __shared__ float sh[32];
  const int t = blockIdx.x;
  float mx = -INFINITY;
  for (int b = threadIdx.x; b < B; b += blockDim.x) mx = fmaxf(mx, logits[(long long)b * T + t]);
  mx = block_reduce(mx, true, sh);
  float s = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) s += expf(logits[(long long)b * T + t] - mx);
  s = block_reduce(s, false, sh);
  const float inv = 1.f / s;
  for (int b = threadIdx.x; b < B; b += blockDim.x) P[(long long)t * B + b] = expf(logits[(long long)b * T + t] - mx) * inv;
}
// dLt[t,b] = P (dP - sum_b P dP);  dlogits[b,t] = dLt[t,b]
__global__ void q_batch_softmax_bwd_kernel(const float* __restrict__ P, const float* __restrict__ dP, int B, int T,
                                           float* __restrict__ dlogits) {
  __shared__ float sh[32];
  const int t = blockIdx.x;
  float s = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) s += P[(long long)t * B + b] * dP[(long long)t * B + b];
  s = block_reduce(s, false, sh);
  for (int b = threadIdx.x; b < B; b += blockDim.x)
    dlogits[(long long)b * T + t] = P[(long long)t * B + b] * (dP[(long long)t * B + b] - s);
}
// language_model.py:165-170: the [T,B] softmax is RAW-reshaped to [B,1,T] (w[b,t] = flat element b*T+t) and multiplies seq.
__global__ void q_pool_fwd_kernel(const float* __restrict__ Wf, const float* __restrict__ seq, int B, int T, int H, float* __restrict__ q_att) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)B * H; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / H), j = (int)(i - (long long)b * H);
    float acc = 0.f;
    for (int t = 0; t < T; ++t) acc += Wf[(long long)b * T + t] * seq[((long long)b * T + t) * H + j];
    q_att[i] = acc;
  }
}
// dseq[b,t,:] = w[b,t] * dq_att[b,:] (+ dq_last[b,:] at t = T-1: q_emb = output[:, -1], language_model.py:120);
// dW[b,t] = <dq_att[b,:], seq[b,t,:]>.  One warp per (b,t).
__global__ void q_pool_bwd_kernel(const float* __restrict__ Wf, const float* __restrict__ seq, const float* __restrict__ dq_att,
                                  const float* __restrict__ dq_last, int B, int T, int H, float* __restrict__ dseq,
                                  float* __restrict__ dW) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarp = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long bt = warp; bt < (long long)B * T; bt += nwarp) {
    const int b = (int)(bt / T), t = (int)(bt - (long long)b * T);
    const float w = Wf[bt];
    float dot = 0.f;
    for (int j = lane; j < H; j += 32) {
      const float g = dq_att[(long long)b * H + j];
      dot += g * seq[bt * H + j];
      float d = w * g;
      if (t == T - 1 && dq_last) d += dq_last[(long long)b * H + j];
      dseq[bt * H + j] = d;
    }
    dot = warp_sum(dot);
    if (lane == 0) dW[bt] = dot;
  }
}

__global__ void q_dot_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float* out) {
  __shared__ float sh[32];
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) s += a[i] * b[i];
  s = block_reduce(s, false, sh);
  if (threadIdx.x == 0) atomicAdd(out, s);
}
// weight_norm.py:41: alpha = g / sqrt(max(sum v^2, 1e-12))
__global__ void q_wn_alpha_kernel(const float* g, const float* vv, float* alpha) { *alpha = *g * rsqrtf(fmaxf(*vv, 1e-12f)); }
// G = dL/dW_eff, W_eff = alpha v:  dv = alpha (G - <G,v> v / ||v||^2),  dg = <G,v> / ||v||
__global__ void q_wn_bwd_kernel(const float* __restrict__ G, const float* __restrict__ v, const float* g, const float* vv,
                                const float* Gv, long long n, float* __restrict__ dv, float* dg) {
  const float nn = fmaxf(*vv, 1e-12f), rs = rsqrtf(nn), alpha = *g * rs, k = *Gv / nn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dv[i] = alpha * (G[i] - k * v[i]);
  if (blockIdx.x == 0 && threadIdx.x == 0) *dg = *Gv * rs;
}
// train.py:112-113: g' = g * clip / max(||g||, clip) per tensor; Keras Adamax with lr_t = lr / (1 - beta1^t)
__global__ void q_clip_adamax_kernel(float* __restrict__ w, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ u,
                                     long long n, const float* gsumsq, float clip, float lr_t, float b1, float b2, float eps) {
  const float scale = clip / fmaxf(sqrtf(*gsumsq), clip);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = grad[i] * scale;
    const float mi = m[i] + (g - m[i]) * (1.f - b1);
    const float ui = fmaxf(b2 * u[i], fabsf(g));
    m[i] = mi; u[i] = ui;
    w[i] -= (lr_t * mi) / (ui + eps);
  }
}

// ---- the embedding tables' gradient is tf.IndexedSlices in the reference (language_model.py:33 reads the variable through
// tf.nn.embedding_lookup), and train.py:112-113 treats it as such:
//   tf.clip_by_norm   takes the norm of the PER-OCCURRENCE values (duplicates not summed):   sumsq = sum_r ||dX[r, col0:col0+E]||^2
//   Adamax (sparse)   m decays everywhere and receives the summed occurrences; u decays everywhere, then every occurrence adds
//                     max(u_row, |value|) - u_row of the decayed row it gathered; the subtraction covers the whole table.
// (Rows of the padding token carry zero values: the mask multiplies the lookup's output.)
__global__ void q_embed_sumsq_kernel(const int* __restrict__ tokens, long long BT, int n_token, int E, int W, int col0,
                                     const float* __restrict__ dX, float* out) {
  float ss = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < BT * E; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / E;
    const int c = (int)(i - r * E);
    const int tok = tokens[r];
    if (tok == n_token || tok < 0 || tok > n_token) continue;
    const float v = dX[r * W + col0 + c];
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0 && ss != 0.f) atomicAdd(out, ss);
}
// uinc[tok, c] += max(beta2 * u[tok, c], |scale * dX[r, col0 + c]|) - beta2 * u[tok, c]   per occurrence r
__global__ void q_embed_uinc_kernel(const int* __restrict__ tokens, long long BT, int n_token, int E, int W, int col0,
                                    const float* __restrict__ dX, const float* __restrict__ u, const float* gsumsq, float clip, float b2,
                                    float* uinc) {
  const float scale = clip / fmaxf(sqrtf(*gsumsq), clip);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < BT * E; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / E;
    const int c = (int)(i - r * E);
    const int tok = tokens[r];
    if (tok == n_token || tok < 0 || tok > n_token) continue;
    const float ud = b2 * u[(long long)tok * E + c];
    const float inc = fmaxf(ud, fabsf(scale * dX[r * W + col0 + c])) - ud;
    if (inc > 0.f) atomicAdd(uinc + (long long)tok * E + c, inc);
  }
}
// dense pass of the sparse Adamax: grad = summed occurrences (scatter-added table), uinc from the kernel above (zeroed on exit)
__global__ void q_clip_adamax_sparse_kernel(float* __restrict__ w, const float* __restrict__ grad, float* __restrict__ m,
                                            float* __restrict__ u, float* __restrict__ uinc, long long n, const float* gsumsq, float clip,
                                            float lr_t, float b1, float b2, float eps) {
  const float scale = clip / fmaxf(sqrtf(*gsumsq), clip);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float g = grad[i] * scale;
    const float mi = m[i] + (g - m[i]) * (1.f - b1);
    const float ui = b2 * u[i] + uinc[i];
    m[i] = mi; u[i] = ui; uinc[i] = 0.f;
    w[i] -= (lr_t * mi) / (ui + eps);
  }
}

int need_device(const char* what) {
  if (regat_device_count() == 0) { set_error("%s: no CUDA device (there is no CPU fallback)", what); return REGAT_ERR_CUDA; }
  return REGAT_OK;
}

}  // namespace
}  // namespace regat

using namespace regat;

extern "C" int regat_q_embed_fwd(const int32_t* tokens, int64_t BT, int n_token, int E, const float* emb, const float* emb2,
                                 float* out, regat_stream_t stream) {
  REGAT_REQUIRE(BT >= 0 && n_token > 0 && E > 0, REGAT_ERR_SHAPE, "q_embed_fwd: bad shape");
  if (BT == 0) return REGAT_OK;
  REGAT_REQUIRE(tokens && emb && out, REGAT_ERR_ARG, "q_embed_fwd: null pointer");
  REGAT_TRY(need_device("q_embed_fwd"));
  q_embed_fwd_kernel<<<grid1d(BT * (emb2 ? 2 * E : E)), 256, 0, (cudaStream_t)stream>>>(tokens, BT, n_token, E, emb, emb2, out);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
extern "C" int regat_q_embed_bwd(const int32_t* tokens, int64_t BT, int n_token, int E, int width, const float* dX, float* demb,
                                 float* demb2, regat_stream_t stream) {
  REGAT_REQUIRE(BT >= 0 && n_token > 0 && E > 0 && (width == E || width == 2 * E), REGAT_ERR_SHAPE, "q_embed_bwd: bad shape");
  if (BT == 0 || (!demb && !demb2)) return REGAT_OK;
  REGAT_REQUIRE(tokens && dX, REGAT_ERR_ARG, "q_embed_bwd: null pointer");
  REGAT_REQUIRE(!demb2 || width == 2 * E, REGAT_ERR_SHAPE, "q_embed_bwd: second table needs width 2E");
  REGAT_TRY(need_device("q_embed_bwd"));
  q_embed_bwd_kernel<<<grid1d(BT * width), 256, 0, (cudaStream_t)stream>>>(tokens, BT, n_token, E, width, dX, demb, demb2);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
extern "C" int regat_q_gru_gates_fwd(int B, int H, const float* xi, int64_t ld_xi, const float* hi, const float* hp, int64_t ld_hp,
                                     float* h_out, int64_t ld_h, float* z, float* r, float* c, float* hp_copy, regat_stream_t stream) {
  REGAT_REQUIRE(B >= 0 && H > 0 && ld_xi >= 3 * (int64_t)H && ld_h >= H && (!hp || ld_hp >= H), REGAT_ERR_SHAPE, "q_gru_gates_fwd: bad shape");
  if (B == 0) return REGAT_OK;
  REGAT_REQUIRE(xi && hi && h_out && z && r && c, REGAT_ERR_ARG, "q_gru_gates_fwd: null pointer");
  REGAT_TRY(need_device("q_gru_gates_fwd"));
  q_gru_gates_fwd_kernel<<<grid1d((long long)B * H), 256, 0, (cudaStream_t)stream>>>(B, H, xi, ld_xi, hi, hp, ld_hp, h_out, ld_h, z, r, c, hp_copy);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
extern "C" int regat_q_gru_gates_bwd(int B, int H, const float* dh_seq, int64_t ld_dseq, const float* dh_rec, const float* z,
                                     const float* r, const float* c, const float* hp_copy, const float* hi, float* dxi,
                                     int64_t ld_dxi, float* dhi, float* dhp, regat_stream_t stream) {
  REGAT_REQUIRE(B >= 0 && H > 0 && ld_dseq >= H && ld_dxi >= 3 * (int64_t)H, REGAT_ERR_SHAPE, "q_gru_gates_bwd: bad shape");
  if (B == 0) return REGAT_OK;
  REGAT_REQUIRE(dh_seq && z && r && c && hp_copy && hi && dxi && dhi && dhp, REGAT_ERR_ARG, "q_gru_gates_bwd: null pointer");
  REGAT_TRY(need_device("q_gru_gates_bwd"));
  q_gru_gates_bwd_kernel<<<grid1d((long long)B * H), 256, 0, (cudaStream_t)stream>>>(B, H, dh_seq, ld_dseq, dh_rec, z, r, c, hp_copy, hi, dxi,
                                                                                   ld_dxi, dhi, dhp);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
extern "C" int regat_q_tanh_fwd(float* x, int64_t n, regat_stream_t stream) {
  REGAT_REQUIRE(n >= 0, REGAT_ERR_SHAPE, "q_tanh_fwd: bad size");
  if (n == 0) return REGAT_OK;
  REGAT_REQUIRE(x, REGAT_ERR_ARG, "q_tanh_fwd: null pointer");
  REGAT_TRY(need_device("q_tanh_fwd"));
  q_tanh_fwd_kernel<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>(x, n);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
extern "C" int regat_q_tanh_bwd(float* dy, const float* y, int64_t n, regat_stream_t stream) {
  REGAT_REQUIRE(n >= 0, REGAT_ERR_SHAPE, "q_tanh_bwd: bad size");
  if (n == 0) return REGAT_OK;
  REGAT_REQUIRE(dy && y, REGAT_ERR_ARG, "q_tanh_bwd: null pointer");
  REGAT_TRY(need_device("q_tanh_bwd"));
  q_tanh_bwd_kernel<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>(dy, y, n);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
extern "C" int regat_q_batch_softmax_fwd(const float* logits, int B, int T, float* P, regat_stream_t stream) {
  REGAT_REQUIRE(B >= 2 && T > 0, REGAT_ERR_SHAPE, "q_batch_softmax_fwd: needs batch >= 2 (tf.squeeze drops a batch of 1, language_model.py:159)");
  REGAT_REQUIRE(logits && P, REGAT_ERR_ARG, "q_batch_softmax_fwd: null pointer");
  REGAT_TRY(need_device("q_batch_softmax_fwd"));
  q_batch_softmax_fwd_kernel<<<T, 256, 0, (cudaStream_t)stream>>>(logits, B, T, P);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
extern "C" int regat_q_batch_softmax_bwd(const float* P, const float* dP, int B, int T, float* dlogits, regat_stream_t stream) {
  REGAT_REQUIRE(B >= 2 && T > 0, REGAT_ERR_SHAPE, "q_batch_softmax_bwd: bad shape");
  REGAT_REQUIRE(P && dP && dlogits, REGAT_ERR_ARG, "q_batch_softmax_bwd: null pointer");
  REGAT_TRY(need_device("q_batch_softmax_bwd"));
  q_batch_softmax_bwd_kernel<<<T, 256, 0, (cudaStream_t)stream>>>(P, dP, B, T, dlogits);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
extern "C" int regat_q_pool_fwd(const float* w_flat, const float* seq, int B, int T, int H, float* q_att, regat_stream_t stream) {
  REGAT_REQUIRE(B > 0 && T > 0 && H > 0, REGAT_ERR_SHAPE, "q_pool_fwd: bad shape");
  REGAT_REQUIRE(w_flat && seq && q_att, REGAT_ERR_ARG, "q_pool_fwd: null pointer");
  REGAT_TRY(need_device("q_pool_fwd"));
  q_pool_fwd_kernel<<<grid1d((long long)B * H), 256, 0, (cudaStream_t)stream>>>(w_flat, seq, B, T, H, q_att);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
extern "C" int regat_q_pool_bwd(const float* w_flat, const float* seq, const float* dq_att, const float* dq_last, int B, int T, int H,
                                float* dseq, float* dW, regat_stream_t stream) {
  REGAT_REQUIRE(B > 0 && T > 0 && H > 0, REGAT_ERR_SHAPE, "q_pool_bwd: bad shape");
  REGAT_REQUIRE(w_flat && seq && dq_att && dseq && dW, REGAT_ERR_ARG, "q_pool_bwd: null pointer");
  REGAT_TRY(need_device("q_pool_bwd"));
  q_pool_bwd_kernel<<<grid1d((long long)B * T * 32), 256, 0, (cudaStream_t)stream>>>(w_flat, seq, dq_att, dq_last, B, T, H, dseq, dW);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
extern "C" int regat_q_dot(const float* a, const float* b, int64_t n, float* out, regat_stream_t stream) {
  REGAT_REQUIRE(n >= 0, REGAT_ERR_SHAPE, "q_dot: bad size");
  if (n == 0) return REGAT_OK;
  REGAT_REQUIRE(a && b && out, REGAT_ERR_ARG, "q_dot: null pointer");
  REGAT_TRY(need_device("q_dot"));
  q_dot_kernel<<<std::min(grid1d(n), 256), 256, 0, (cudaStream_t)stream>>>(a, b, n, out);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
extern "C" int regat_q_wn_alpha(const float* g, const float* vv, float* alpha, regat_stream_t stream) {
  REGAT_REQUIRE(g && vv && alpha, REGAT_ERR_ARG, "q_wn_alpha: null pointer");
  REGAT_TRY(need_device("q_wn_alpha"));
  q_wn_alpha_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(g, vv, alpha);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
extern "C" int regat_q_wn_bwd(const float* G, const float* v, const float* g, const float* vv, const float* Gv, int64_t n, float* dv,
                              float* dg, regat_stream_t stream) {
  REGAT_REQUIRE(n > 0, REGAT_ERR_SHAPE, "q_wn_bwd: bad size");
  REGAT_REQUIRE(G && v && g && vv && Gv && dv && dg, REGAT_ERR_ARG, "q_wn_bwd: null pointer");
  REGAT_TRY(need_device("q_wn_bwd"));
  q_wn_bwd_kernel<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>(G, v, g, vv, Gv, n, dv, dg);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
extern "C" int regat_q_clip_adamax(float* w, const float* grad, float* m, float* u, int64_t n, const float* gsumsq, float clip,
                                   float lr, int step, float beta1, float beta2, float eps, regat_stream_t stream) {
  REGAT_REQUIRE(n > 0 && step >= 1, REGAT_ERR_SHAPE, "q_clip_adamax: bad size or step");
  REGAT_REQUIRE(w && grad && m && u && gsumsq, REGAT_ERR_ARG, "q_clip_adamax: null pointer");
  REGAT_TRY(need_device("q_clip_adamax"));
  const float lr_t = (float)((double)lr / (1.0 - pow((double)beta1, (double)step)));
  q_clip_adamax_kernel<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>(w, grad, m, u, n, gsumsq, clip, lr_t, beta1, beta2, eps);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}

extern "C" int regat_q_embed_sumsq(const int32_t* tokens, int64_t BT, int n_token, int E, int width, int col0, const float* dX, float* out,
                                   regat_stream_t stream) {
  REGAT_REQUIRE(BT >= 0 && n_token > 0 && E > 0 && col0 >= 0 && col0 + E <= width, REGAT_ERR_SHAPE, "q_embed_sumsq: bad shape");
  if (BT == 0) return REGAT_OK;
  REGAT_REQUIRE(tokens && dX && out, REGAT_ERR_ARG, "q_embed_sumsq: null pointer");
  REGAT_TRY(need_device("q_embed_sumsq"));
  q_embed_sumsq_kernel<<<grid1d(BT * E), 256, 0, (cudaStream_t)stream>>>(tokens, BT, n_token, E, width, col0, dX, out);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
extern "C" int regat_q_embed_clip_adamax(const int32_t* tokens, int64_t BT, int n_token, int E, int width, int col0, const float* dX,
                                         float* table, const float* grad_dense, float* m, float* u, float* uinc, const float* gsumsq,
                                         float clip, float lr, int step, float beta1, float beta2, float eps, regat_stream_t stream) {
  REGAT_REQUIRE(BT >= 0 && n_token > 0 && E > 0 && col0 >= 0 && col0 + E <= width && step >= 1, REGAT_ERR_SHAPE, "q_embed_clip_adamax: bad shape");
  REGAT_REQUIRE(tokens && dX && table && grad_dense && m && u && uinc && gsumsq, REGAT_ERR_ARG, "q_embed_clip_adamax: null pointer");
  REGAT_TRY(need_device("q_embed_clip_adamax"));
  cudaStream_t st = (cudaStream_t)stream;
  if (BT > 0) {
    q_embed_uinc_kernel<<<grid1d(BT * E), 256, 0, st>>>(tokens, BT, n_token, E, width, col0, dX, u, gsumsq, clip, beta2, uinc);
    REGAT_POST_LAUNCH();
  }
  const long long n = (long long)(n_token + 1) * E;
  const float lr_t = (float)((double)lr / (1.0 - pow((double)beta1, (double)step)));
  q_clip_adamax_sparse_kernel<<<grid1d(n), 256, 0, st>>>(table, grad_dense, m, u, uinc, n, gsumsq, clip, lr_t, beta1, beta2, eps);
  REGAT_POST_LAUNCH();
  return REGAT_OK;
}
