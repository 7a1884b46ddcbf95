/* regat.h -- C ABI of libregat.so: the B200 (sm_100a) implementation of ReGAT's
 * implicit-relation graph-attention + BUTD-fusion hot path.
 *
 * The reference (jhss/TF_VQA_ReGAT) has no FFI; its boundary for this path is the tf.keras
 * layer API.  Each entry point below names the reference code it stands in for
 * (file:line under the reference tree).  INTEGRATION.md shows the ctypes / TF-DLPack
 * binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; row-major, contiguous;
 *   - `dtype` is the storage type of activations: REGAT_F32 or REGAT_BF16 (fp32 accumulate);
 *     parameters, optimizer state, weight gradients, logits and the loss are always fp32;
 *   - tensors are BORROWED for the duration of the call; the library never allocates or
 *     frees caller memory and keeps no reference to it (engine_bind excepted: it keeps
 *     the pointers until the next bind/destroy);
 *   - every call is asynchronous on `stream` (a cudaStream_t) and performs no hidden sync;
 *   - return value: 0 = ok, negative = regat_status; message via regat_last_error()
 *     (thread-local).  No exception or abort crosses this boundary;
 *   - there is NO CPU fallback: without a CUDA device every compute entry returns
 *     REGAT_ERR_CUDA.
 */
#ifndef REGAT_H_
#define REGAT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REGAT_ABI_VERSION 1

#if defined(__GNUC__)
#define REGAT_API __attribute__((visibility("default")))
#else
#define REGAT_API
#endif

typedef void* regat_stream_t; /* cudaStream_t */

typedef enum regat_status {
  REGAT_OK = 0,
  REGAT_ERR_ARG = -1,         /* null pointer / bad enum */
  REGAT_ERR_SHAPE = -2,       /* unsupported or inconsistent shape */
  REGAT_ERR_DTYPE = -3,
  REGAT_ERR_DEVICE = -4,      /* DLPack tensor not on the current CUDA device */
  REGAT_ERR_ALIGN = -5,       /* pointer / leading dimension not 16-byte aligned */
  REGAT_ERR_CUDA = -6,        /* CUDA runtime / driver error (incl. no device) */
  REGAT_ERR_UNSUPPORTED = -7,
  REGAT_ERR_WORKSPACE = -8    /* caller-provided buffer too small */
} regat_status;

typedef enum regat_dtype { REGAT_F32 = 0, REGAT_BF16 = 1 } regat_dtype;

/* Hyper-parameters of the path.  Defaults = reference config/butd_vqa.json:1-29. */
typedef struct regat_config {
  int32_t v_dim;        /* 2048  region feature dim                                   */
  int32_t q_dim;        /* 768   num_hid: question dim = BUTD hidden                  */
  int32_t rel_dim;      /* 1024  relation_dim                                         */
  int32_t num_heads;    /* 16    (head dim rel_dim/num_heads must be 64)              */
  int32_t pos_emb_dim;  /* 64    imp_pos_emb_dim                                      */
  int32_t nongt_dim;    /* 20                                                        */
  int32_t dir_num;      /* 2                                                         */
  int32_t num_answers;  /* 3129                                                      */
  int32_t label_bias;   /* 0     graph_att_net.py:25 use_bias of the label FC         */
  int32_t residual;     /* 1     relation_encoder.py:88-91                            */
  float grad_clip;      /* 0.25  per-tensor clip_by_norm, train.py:112 / main.py:24   */
  float beta1, beta2, eps; /* Adamax 0.9, 0.999, 1e-8, train.py:48                    */
} regat_config;

REGAT_API int regat_abi_version(void);
/* Copies the calling thread's last error message (NUL-terminated) into buf; returns its length. */
REGAT_API int regat_last_error(char* buf, size_t n);
/* Fills cfg with the reference defaults above. */
REGAT_API int regat_default_config(regat_config* cfg);
/* Number of CUDA devices visible (0 on a CPU-only box); never fails. */
REGAT_API int regat_device_count(void);
/* For bindings whose tensor provider exposes neither copies nor a stream (TensorFlow eager): kind 1 = host to
 * device, 2 = device to host (both return after the copy has completed), 3 = device to device (asynchronous on `stream`);
 * regat_device_synchronize waits for all work on the current device. */
REGAT_API int regat_memcpy(void* dst, const void* src, int64_t bytes, int kind, regat_stream_t stream);
REGAT_API int regat_device_synchronize(void);

/* ------------------------------------------------------------------ stage 1 ----------
 * Materialised position embedding, the drop-in for
 * position_emb.py:153-160 prepare_graph_variables (-> :117-151, :96-115):
 *   boxes [B,N,4] fp32 abs-pixel (x1,y1,x2,y2)  ->  pos_emb [B,M,N,feat_dim] fp32,
 *   M = min(nongt_dim, N).  wave_div_host: feat_dim/8 fp32 divisors 1000^(8k/feat_dim)
 *   computed by the HOST binding in fp32 exactly as position_emb.py:98-100 does.
 * The fused attention kernel never needs this tensor; it exists for callers that want the
 * reference's array (compat path) and for parity tests of the geometry code.            */
REGAT_API int regat_position_embedding(const float* boxes, int B, int N, int nongt_dim, int feat_dim,
                             const float* wave_div_host, float* pos_emb, regat_stream_t stream);

/* ------------------------------------------------------------------ weight norm -------
 * weight_norm.py:35-41: W = g * v / sqrt(max(sum v^2, 1e-12)), scalar g, whole-tensor norm.
 * Because the norm is a scalar, W = alpha*v; GEMMs read v and scale by alpha in their epilogue.
 *   sumsq[l] += sum(v_l^2) for each of n_layers tensors (caller zeroes sumsq);
 *   v_lowp (optional, bf16): a bf16 copy of each v with leading dimension ld_lowp[l]
 *   (>= cols, multiple of 8) at element offset off_lowp[l].
 * descriptors are HOST arrays of length n_layers.                                       */
REGAT_API int regat_wn_prepare(const float* params, const int64_t* v_off_host, const int64_t* v_numel_host,
                     const int32_t* v_cols_host, int n_layers, float* sumsq,
                     void* v_lowp, const int64_t* off_lowp_host, const int32_t* ld_lowp_host,
                     regat_stream_t stream);
/* alpha[l] = g_l * rsqrt(max(sumsq[l], 1e-12)); inv_norm[l] = rsqrt(max(sumsq[l],1e-12)). */
REGAT_API int regat_wn_alpha(const float* params, const int64_t* g_off_host, int n_layers, const float* sumsq,
                   float* alpha, float* inv_norm, regat_stream_t stream);

/* ------------------------------------------------------------------ dense (fc.py:11-50) --
 * C[M,N] = epilogue( op(A)[M,K] . op(B)[K,N] ), fp32 accumulate.
 *   transA = 0: A stored [M,K] (lda >= K);  1: A stored [K,M] (lda >= M)
 *   transB = 0: B stored [K,N] (ldb >= N);  1: B stored [N,K] (ldb >= K)
 * REGAT_F32 runs an exact-fp32 SIMT kernel (parity mode); REGAT_BF16 runs the TMA + tcgen05
 * kernel (A, B bf16).  Epilogue, applied per element x = acc in this order:
 *   x += row_scale[r] * addend[(r / addend_rows) * addend_ld + c]   (if addend)
 *   x *= alpha[c / alpha_cols]    (alpha_cols == 0: alpha[0]; alpha == NULL: 1)
 *   x += bias[c]                  (if bias)
 *   x  = max(x, 0)                (if relu)
 *   x += C_old[r,c]               (if accumulate)
 *   x  = gate[r,c] > 0 ? x : 0    (if gate; same dtype/ld as given)
 *   C[r,c] = x                    (dtype c_dtype: REGAT_F32 or the activation dtype)
 *   C2[(r / c2_rows_in) * c2_rows_keep + r % c2_rows_in, c] = x  if r % c2_rows_in < c2_rows_keep
 * split_k > 1 (bf16 path, c_dtype F32, no epilogue but alpha): the K range is split over CTAs, C is
 * zeroed by the call and partial sums are atomically added (used for weight gradients).   */
typedef struct regat_epilogue {
  const float* alpha; int32_t alpha_cols;
  const float* bias;
  const float* addend; int32_t addend_ld; int32_t addend_rows; const float* row_scale;
  int32_t relu;
  int32_t accumulate;
  const void* gate; int32_t gate_ld;
  void* c2; int32_t c2_ld; int32_t c2_rows_in; int32_t c2_rows_keep;
  int32_t split_k;
} regat_epilogue;

REGAT_API int regat_gemm(int dtype, int transA, int transB, int M, int N, int K,
               const void* A, int lda, const void* B, int ldb,
               void* C, int ldc, int c_dtype, const regat_epilogue* epi, regat_stream_t stream);

/* ------------------------------------------------------------------ kernel (a) ---------
 * Fused geometry-bias graph attention, forward.  Replaces, for all dir_num directions and all
 * heads of one GraphAttentionNetwork call: position_emb.py:96-151 (geometry + sinusoid),
 * graph_att_layer.py:63-111 (logits, pair_pos_fc bias, relu/max/log, adjacency where (a no-op
 * for the implicit relation, relation_encoder.py:76), label bias, softmax over the first
 * M = min(nongt_dim, N) objects, aggregation), the grouped 1x1 conv graph_att_layer.py:112-117
 * (re-associated into V' = s[:, :M] . Kc + bc, computed by regat_gemm beforehand), the sum
 * over directions + ReLU graph_att_net.py:64-81, and the residual relation_encoder.py:88-91.
 *
 *   q     [B*N, dirs*D]   : Q_d at column d*D                         (activation dtype)
 *   kv    [B*M, 2*dirs*D] : K_d at column d*D, V'_d at (dirs+d)*D
 *   boxes [B,N,4] fp32, or pos_emb [B,M,N,E] fp32 (exactly one non-NULL)
 *   wg    : dirs pointers' worth of pair_pos_fc: v [dirs][E,H] fp32 raw kernels (contiguous per
 *           dir at wg + d*wg_stride), alpha_g [dirs], bg [dirs][H] (at bg + d*bg_stride)
 *   label_c : device scalar c = WN(label FC)(1) (+bias)  (graph_att_net.py:71)
 *   s, v0 [B*N, D] ; out v1 [B*N, D] = (residual ? v0 : 0) + relu(s + sum_d O_d);
 *   s == NULL: v1 = sum_d O_d, the raw output of GraphSelfAttentionLayer.call (graph_att_layer.py:121)
 *   save_p, save_gbias [B,dirs,H,N,M] fp32 and gate [B*N, H] uint64 (bit e of word (row,h) set
 *   iff relu input > 0) are written when non-NULL (training).
 * Head dim must be 64; N <= 128; 16*M*dirs*H*4 bytes must fit in shared memory.          */
REGAT_API int regat_geoattn_fwd(int dtype, int B, int N, int nongt_dim, int D, int H, int dirs, int E,
                      const void* q, const void* kv, const float* boxes, const float* pos_emb,
                      const float* wave_div_host,
                      const float* wg, int64_t wg_stride, const float* alpha_g,
                      const float* bg, int64_t bg_stride, const float* label_c,
                      const void* s, const void* v0, int residual, void* v1,
                      float* save_p, float* save_gbias, uint64_t* gate, regat_stream_t stream);

/* Backward of the attention part.  Consumes dv1 [B*N,D] (gradient w.r.t. the encoder output),
 * the saved P / gbias / gate, Q and KV; produces
 *   dq [B*N, dirs*D], dkv [B*M, 2*dirs*D]  (activation dtype),
 *   dout [B*N, D] = dv1 * gate  (the gradient of relu's input: the direct `s` term),
 *   and overwrites save_p in place with dL (gradient w.r.t. the logits), which
 *   regat_geo_bwd then reduces.                                                           */
REGAT_API int regat_attn_bwd(int dtype, int B, int N, int nongt_dim, int D, int H, int dirs,
                   const void* q, const void* kv, const void* dv1, const uint64_t* gate,
                   float* p_inout_dl, void* dq, void* dkv, void* dout, regat_stream_t stream);

/* Backward of the geometry bias: recomputes Emb(i',j') from the boxes (never stored) and reduces
 *   dWg[d][e][h] += sum Emb * dz,  dbg[d][h] += sum dz,  dc += sum dL
 * with dz = dL / z where z = exp(gbias) >= 1e-6 (graph_att_layer.py:79-88), else 0.
 * dwg receives the gradient w.r.t. the EFFECTIVE kernel W = alpha*v (same layout as wg).  */
REGAT_API int regat_geo_bwd(int B, int N, int nongt_dim, int H, int dirs, int E,
                  const float* boxes, const float* pos_emb, const float* wave_div_host,
                  const float* dl, const float* gbias,
                  float* dwg, int64_t dwg_stride, float* dbg, int64_t dbg_stride, float* dc,
                  regat_stream_t stream);
/* Same reduction; fast_math != 0 (bf16 training mode, boxes path) evaluates sin/cos/log/exp on the SFU and the
 * 64 x (dirs*H) outer products in one TF32 pass instead of 3xTF32 -- the precision class of the bf16 forward path. */
REGAT_API int regat_geo_bwd_ex(int B, int N, int nongt_dim, int H, int dirs, int E,
                  const float* boxes, const float* pos_emb, const float* wave_div_host,
                  const float* dl, const float* gbias,
                  float* dwg, int64_t dwg_stride, float* dbg, int64_t dbg_stride, float* dc,
                  int fast_math, regat_stream_t stream);

/* bf16 training fast path of the three calls above (boxes given, M = min(nongt_dim, N) <= 24 keys -- every BASELINE config;
 * regat_geoattn_fast_supported says whether a shape qualifies).  Same math, sized for bytes and instructions: SFU sin/cos/log/exp,
 * one-pass TF32 geometry projection, bf16 mma for Q K^T and P V'.  What is kept for the backward pass is packed bf16
 * [B][dirs*H][N][MPAD], MPAD = M rounded up to even:
 *   save_p16  the probabilities (the very bf16 pairs that fed the P V' mma);
 *   save_rz16 rz = 1/z where the pair_pos_fc pre-activation z is above the relu / 1e-6 clamp, else 0 (graph_att_layer.py:79-88).
 * regat_attn_bwd_fast overwrites save_p16 in place with dz = dL * rz (dL = gradient w.r.t. the logits) and adds sum(dL) to *dc
 * (the label constant's gradient, optional); regat_geo_bwd_fast reduces dz into dWg / dbg with recomputed embeddings.
 * q, kv, s, v0, v1, dv1, dq, dkv, dout are bf16 with the layouts documented above. */
REGAT_API int regat_geoattn_fast_supported(int N, int nongt_dim);
REGAT_API int regat_geoattn_fwd_fast(int B, int N, int nongt_dim, int D, int H, int dirs, int E, const void* q, const void* kv,
                           const float* boxes, const float* wave_div_host, const float* wg, int64_t wg_stride,
                           const float* alpha_g, const float* bg, int64_t bg_stride, const float* label_c, const void* s,
                           const void* v0, int residual, void* v1, void* save_p16, void* save_rz16, uint64_t* gate,
                           regat_stream_t stream);
REGAT_API int regat_attn_bwd_fast(int B, int N, int nongt_dim, int D, int H, int dirs, const void* q, const void* kv,
                        const void* dv1, const uint64_t* gate, void* p16_inout_dz, const void* rz16, void* dq, void* dkv,
                        void* dout, float* dc, regat_stream_t stream);
REGAT_API int regat_geo_bwd_fast(int B, int N, int nongt_dim, int H, int dirs, int E, const float* boxes,
                       const float* wave_div_host, const void* dz16, float* dwg, int64_t dwg_stride, float* dbg,
                       int64_t dbg_stride, regat_stream_t stream);

/* ------------------------------------------------------------------ explicit relation ---------------------
 * ExplicitRelationEncoder (relation_encoder.py:95-143) = the same GraphAttentionNetwork with pos_emb_dim = -1 (no pair_pos_fc,
 * no geometry) and a labelled adjacency adj [B,N,N,L] (one-hot edge labels; spatial L = 11, semantic L = 15).  Per direction d
 * (d = 1 uses adj transposed in its two object axes, graph_att_net.py:56) and restricted to the first M = min(nongt_dim, N) keys:
 *     label_att = bias-FC(adj_d) = sum_l adj_d[..,l] * w_label[l] (+ b_label)          graph_att_net.py:71
 *     logits    = where(sum_l adj_d > 0, Q K^T / 8, -9e15) + label_att                 graph_att_layer.py:90-102
 * regat_explicit_pair_bias builds pair_bias [B,dirs,N,M] = label_att where the adjacency is set, -9e15 where it is not (in fp32
 * -9e15 absorbs both the affinity and the label term, exactly as the reference's where + add does); w_label is the EFFECTIVE
 * kernel alpha*v of the label FC [L,1].  regat_graphattn_explicit_fwd / _bwd are regat_geoattn_fwd / regat_attn_bwd with that
 * term in place of the geometry bias (same layouts; P / dL fp32 [B,dirs,H,N,M]); masked pairs pass no gradient to Q and K,
 * while dL still reaches the label bias: regat_explicit_pair_bias_bwd adds dw_label[l] += sum adj_d * (sum_h dL) and
 * db_label += sum dL (caller zeroes both).                                                                              */
REGAT_API int regat_explicit_pair_bias(int B, int N, int nongt_dim, int L, int dirs, const float* adj, const float* w_label,
                             const float* b_label, float* pair_bias, regat_stream_t stream);
REGAT_API int regat_explicit_pair_bias_bwd(int B, int N, int nongt_dim, int L, int dirs, int H, const float* adj, const float* dl,
                                 float* dw_label, float* db_label, regat_stream_t stream);
REGAT_API int regat_graphattn_explicit_fwd(int dtype, int B, int N, int nongt_dim, int D, int H, int dirs, const void* q,
                                 const void* kv, const float* pair_bias, const void* s, const void* v0, int residual, void* v1,
                                 float* save_p, uint64_t* gate, regat_stream_t stream);
REGAT_API int regat_graphattn_explicit_bwd(int dtype, int B, int N, int nongt_dim, int D, int H, int dirs, const void* q,
                                 const void* kv, const void* dv1, const uint64_t* gate, const float* pair_bias,
                                 float* p_inout_dl, void* dq, void* dkv, void* dout, regat_stream_t stream);

/* ------------------------------------------------------------------ BUTD pooling --------
 * fusion.py:43-54 + :34 with the (linear) v2attention FC re-associated:
 *   logit[b,n] = <v1[b,n,:], weff[b,:]> + cb[b];  att = softmax_n;  pooled[b,:] = sum_n att*v1.
 * weff [B,D] and cb [B] are produced by regat_gemm / regat_rowdot from q2attention and linear. */
REGAT_API int regat_butd_pool_fwd(int dtype, int B, int N, int D, const void* v1, const void* weff,
                        const float* cb, float* att, void* pooled, regat_stream_t stream);
/* dv1[b,n,:] = att*dpooled + dlogit*weff ; dweff[b,:] = sum_n dlogit*v1 ; dcb[b] = sum_n dlogit */
REGAT_API int regat_butd_pool_bwd(int dtype, int B, int N, int D, const void* v1, const void* weff,
                        const float* att, const void* dpooled, void* dv1, void* dweff, float* dcb,
                        regat_stream_t stream);

/* Device half of the reference's batch assembly (dataset.py:329-346: keras pad_sequences(padding='post', maxlen = longest
 * sample)): a batch shipped RAGGED -- packed[total_rows, width] holds the samples' rows back to back, offsets[B+1] (int32,
 * device) the first row of each sample -- becomes padded[B, N, width] with zero rows after each sample's own, exactly what
 * the host-side collate produces, without moving the padding over PCIe.  width must be a multiple of 4 floats, buffers
 * 16-byte aligned.  Offsets live on the device, so the call cannot return an error for them: a sample whose offsets are
 * invalid (negative or decreasing, more than N rows, rows past total_rows) is never read -- it comes out all zero -- and,
 * if `bad` (one int32 on the device, zeroed by the caller) is given, `bad` receives 1 + the index of such a sample. */
REGAT_API int regat_pad_ragged(int B, int N, int width, int64_t total_rows, const float* packed, const int32_t* offsets,
                     float* padded, int32_t* bad, regat_stream_t stream);

/* Elementwise fp32 <-> bf16 conversion of n (multiple of 8) elements -- used for the reduced-precision gradient exchange of the
 * data-parallel path and by hosts that hold fp32 tensors. */
REGAT_API int regat_cast(int from_dtype, int to_dtype, const void* in, void* out, int64_t n, regat_stream_t stream);

/* Data-parallel gradient exchange over NVLink / NVSwitch peer memory (the reference is single-GPU, train.py:111
 * produces the gradients this sums over replicas).  The caller owns two SYMMETRIC allocations, identical on every rank and
 * mapped into every peer: a bf16 staging buffer as long as the flat gradient buffer, and a zero-initialised flag buffer of
 * 2 x 16 uint32.  stage_ptrs / flag_ptrs are HOST arrays of `world` device addresses (entry r = rank r's copy as mapped into
 * this process); multicast_ptr is the NVSwitch multicast address of the staging buffer, or 0 (then peers are read and
 * written one by one).  Per range of the flat buffer, on one stream and in the same order on every rank:
 *     regat_cast(fp32 -> bf16, grads + offset -> local staging + offset)
 *     regat_dp_reduce_bcast(...)     rank r sums slice r of the range over all ranks and writes the sum to all of them
 *     regat_dp_wait_unpack(...)      waits for every slice, then staging + offset -> dst (fp32, dst = grads + offset)
 * epoch: a counter the caller increments per range exchanged (same sequence on every rank, starting at 1).  offset and numel
 * in elements, multiples of 8.  blocks: CTAs of the reduce kernel (<= 0: default 32); few, so it runs beside compute. */
REGAT_API int regat_dp_reduce_bcast(const uint64_t* stage_ptrs, uint64_t multicast_ptr, const uint64_t* flag_ptrs,
                  int rank, int world, int64_t offset, int64_t numel, uint32_t epoch, int blocks, regat_stream_t stream);
/* fp32 wire format, in place: the flat gradient buffer ITSELF is the symmetric allocation (grad_ptrs[r] = rank r's copy,
 * multicast_ptr its multicast address or 0).  One call sums [offset, offset+numel) over all ranks into every rank's buffer
 * (reduce/broadcast kernel + a one-block completion wait); no staging, no pack / unpack passes.  Multiples of 4 elements. */
REGAT_API int regat_dp_allreduce_f32(const uint64_t* grad_ptrs, uint64_t multicast_ptr, const uint64_t* flag_ptrs,
                  int rank, int world, int64_t offset, int64_t numel, uint32_t epoch, int blocks, regat_stream_t stream);
/* Same, with the call counter in device memory (*epoch_dev, zero-initialised, ordinary device memory of this rank): each call
 * uses *epoch_dev + 1 and stores it back when the range is complete, so the launch takes no host scalar that changes from
 * call to call and can sit inside a replayed CUDA graph.  Uses its own flag buffer (do not share one with the host-epoch calls). */
REGAT_API int regat_dp_allreduce_f32_dev(const uint64_t* grad_ptrs, uint64_t multicast_ptr, const uint64_t* flag_ptrs,
                  int rank, int world, int64_t offset, int64_t numel, uint32_t* epoch_dev, int blocks, regat_stream_t stream);
REGAT_API int regat_dp_wait_unpack(const void* stage_local, float* dst, const uint64_t* flag_ptrs, int rank, int world,
                  int64_t offset, int64_t numel, uint32_t epoch, regat_stream_t stream);

/* relation_encoder.py:13-37 concat_visual_question(q, v, mask=True):
 *   mask[b,n] = (sum_d v[b,n,d] != 0) (optional output, fp32);  out[b,n,:] = [ v[b,n,:] || mask * q[b,:] ]   [B,N,D+Q] */
REGAT_API int regat_concat_visual_question(int dtype, int B, int N, int D, int Q, const void* v, const void* q,
                                 void* out, float* mask, regat_stream_t stream);
/* fusion.py:47-52 re-associated: uw[b,c] = u[b,c] * alpha_linear * v_linear[c];  cb[b] = sum_c bias_v2att[c]*uw[b,c]
 * + bias_linear.  u = q2attention(question) with row pitch ldu. */
REGAT_API int regat_butd_prep(int dtype, int B, int Hd, const void* u, int ldu, const float* v_linear,
                    const float* alpha_linear, const float* bias_v2att, const float* bias_linear, void* uw,
                    float* cb, regat_stream_t stream);
/* out = a * b elementwise (fusion.py:39 joint_emb = weighted_visual * question). */
REGAT_API int regat_mul(int dtype, int rows, int cols, const void* a, int lda, const void* b, int ldb, void* out, int ldo,
              regat_stream_t stream);

/* ------------------------------------------------------------------ loss (train.py:20-26,107-108)
 * loss = A * mean_{b,a} BCEwithlogits(logits, target) (added into *loss, caller zeroes);
 * dlogits = (sigmoid(x) - z) / B, written with leading dim ld_d in dtype d_dtype (padding
 * columns zeroed).  score (optional) += sum_b target[b, argmax_a logits[b,:]]  (train.py:28-39). */
REGAT_API int regat_bce_fwd_bwd(int B, int A, const float* logits, int ld_logits, const float* target,
                      float* loss, float* score, void* dlogits, int ld_d, int d_dtype,
                      regat_stream_t stream);

/* ------------------------------------------------------------------ engine --------------
 * The whole hot path of rel_graph_net.py:53-62 (+ train.py:103-113) as one call sequence on
 * one stream with static buffers (CUDA-graph capturable).                                 */
typedef struct regat_engine regat_engine;

REGAT_API int regat_engine_create(const regat_config* cfg, int dtype, int max_batch, int max_rois,
                        regat_engine** out);
REGAT_API int regat_engine_destroy(regat_engine* e);
/* Element counts of the flat fp32 parameter buffer (same for grads and each Adamax slot) and
 * byte size of the workspace.  Layout = Keras variable order, see tf_vqa_regat_b200/config.py. */
REGAT_API int regat_engine_sizes(const regat_engine* e, int64_t* param_elems, int64_t* workspace_bytes);
/* Offset (elements) and size of parameter tensor idx in layout order; returns REGAT_ERR_ARG past the end. */
REGAT_API int regat_engine_param(const regat_engine* e, int idx, int64_t* offset, int64_t* numel,
                       int32_t* layer, int32_t* kind /*0=v,1=g,2=bias*/);
REGAT_API int regat_engine_bind(regat_engine* e, float* params, float* grads, float* adamax_m,
                      float* adamax_u, void* workspace, int64_t workspace_bytes);
/* Forward only (eval, train.py:136-177): logits [B, num_answers] fp32, optional att [B,N].
 * features [B,N,v_dim] fp32, boxes [B,N,4] fp32, q_att/q_last [B,q_dim] fp32.             */
REGAT_API int regat_engine_forward(regat_engine* e, int B, int N, const float* features, const float* boxes,
                         const float* q_att, const float* q_last, float* logits, float* att,
                         regat_stream_t stream);
/* Forward + backward: fills grads with dL/dW_eff for every kernel (W_eff = g*v/||v||) and dL/db for every
 * bias -- linear in the batch, hence the quantity to all-reduce (see regat_engine_finalize_grads);
 * loss_out[0] = loss, loss_out[1] = batch score; dq_att/dq_last [B,q_dim] fp32 (optional) for the
 * upstream language model.  grad_scale multiplies the loss gradient (1/world for data parallel). */
REGAT_API int regat_engine_fwd_bwd(regat_engine* e, int B, int N, const float* features, const float* boxes,
                         const float* q_att, const float* q_last, const float* target,
                         float grad_scale, float* loss_out, float* logits, float* dq_att,
                         float* dq_last, regat_stream_t stream);
/* Per-tensor clip_by_norm + Adamax (train.py:112-113) on the bound grads; step is 1-based. */
REGAT_API int regat_engine_update(regat_engine* e, float lr, int step, regat_stream_t stream);
/* fwd_bwd + update (single GPU). */
REGAT_API int regat_engine_train_step(regat_engine* e, int B, int N, const float* features,
                            const float* boxes, const float* q_att, const float* q_last,
                            const float* target, float lr, int step, float* loss_out,
                            regat_stream_t stream);
/* Device-resident optimizer scalars.  regat_engine_train_step_dev runs one whole step (forward, backward, per-tensor clip,
 * Adamax, train.py:103-113) without any host scalar in a launch: the learning rate and the number of steps taken live in the
 * workspace (set_lr / set_step write them with a one-thread kernel on `stream`; the step itself increments the counter and
 * derives lr_t = lr / (1 - beta1^t)).  The call is therefore capturable in ONE CUDA graph and replayable; clip + Adamax of each
 * gradient range run on an internal stream behind the rest of the backward pass, and everything the next forward pass derives
 * from the parameters (alpha = g/||v||, bf16 kernels, gathered biases) is rebuilt there too.
 * regat_engine_update / regat_engine_train_step (host lr, 1-based step) are the same kernels after writing those scalars.
 * regat_engine_get_step synchronises `stream` (tests / checkpoints only).                                                  */
REGAT_API int regat_engine_set_lr(regat_engine* e, float lr, regat_stream_t stream);
REGAT_API int regat_engine_set_step(regat_engine* e, int steps_done, regat_stream_t stream);
REGAT_API int regat_engine_get_step(regat_engine* e, int* steps_done, float* lr, regat_stream_t stream);
REGAT_API int regat_engine_train_step_dev(regat_engine* e, int B, int N, const float* features,
                                const float* boxes, const float* q_att, const float* q_last,
                                const float* target, float* loss_out, regat_stream_t stream);
/* Data parallel inside the engine (train.py:111-113 on every replica after the exchange).  The bound grads buffer is
 * this rank's part of a SYMMETRIC allocation: grad_ptrs[r] / flag_ptrs[r] are rank r's mappings of the gradient buffer and of a
 * zero-initialised flag buffer of >= 32 uint32 (multicast_ptr: the NVSwitch multicast mapping of the gradient buffer, 0 if
 * none).  regat_engine_train_step[_dev] then uses grad_scale = 1/world and reduces each gradient range in place
 * (csrc/dp_exchange.cu) as soon as the backward pass has finished it, before that range's clip + Adamax.  Every rank must
 * issue the same sequence of steps.  world <= 1 switches it off.                                                            */
REGAT_API int regat_engine_set_dp(regat_engine* e, const uint64_t* grad_ptrs, uint64_t multicast_ptr,
                        const uint64_t* flag_ptrs, int rank, int world, int blocks);
/* Re-derives alpha, the bf16 kernels, the gathered biases and the label constant from the parameter buffer on `stream`, now.
 * Needed only when the parameters were written from outside AND the next engine work is the replay of an already captured
 * graph (eager calls do it themselves after regat_engine_params_changed / regat_engine_bind).                               */
REGAT_API int regat_engine_refresh_weights(regat_engine* e, regat_stream_t stream);
/* Data parallel: fn(user, offset, numel) is called from inside regat_engine_fwd_bwd (on the calling thread, while the
 * call is still enqueueing) each time a contiguous range of the grads buffer has received its last write on `stream`:
 * first the BUTD + classifier tail, then self_weights + attention layers, then v2out.  The callee typically records an
 * event on `stream` and starts an all-reduce of that range on another stream, overlapping the rest of the backward.  */
typedef void (*regat_grad_ready_fn)(void* user, int64_t offset, int64_t numel);
REGAT_API int regat_engine_set_grad_callback(regat_engine* e, regat_grad_ready_fn fn, void* user);
/* Measurement aid (tools/gemm_trace.py): with a device buffer of >= 256 int64 the tcgen05 GEMM launches that follow record
 * clock64 stamps of CTA 0 -- [0] kernel start, [1] roles done, [2] exit, then per unit u (u < 14) at [8 + 8u + k]:
 * k=0/1 first/last TMA issue, 2 accumulator stage free, 3 first operands landed, 4 last MMA issued, 5 epilogue ready,
 * 6 accumulator complete, 7 unit stored.  NULL switches it off (the default).  Not thread-safe; never leave it on. */
REGAT_API int regat_gemm_trace(void* device_buf);
/* Measurement aid (bench.py's GEMM-class figure): with on != 0 every dense product of the following EAGER engine calls is
 * bracketed by CUDA timing events; profile_read synchronises the device, returns up to max_records entries -- mnk[3i..3i+2] =
 * (M, N, K), ms[i] -- in launch order and clears the list.  Never enable it while a CUDA graph is being captured. */
REGAT_API int regat_engine_profile(regat_engine* e, int on);
REGAT_API int regat_engine_profile_read(regat_engine* e, int max_records, int32_t* mnk, float* ms, int* count);
/* Number of kernel launches the last engine call issued (for bench.py's gpu_launches). */
REGAT_API int regat_engine_last_launches(const regat_engine* e);
/* Tell the engine that the caller wrote the parameter buffer (checkpoint load, set_weights).  The engine keeps what it
 * derives from the parameters (per-tensor ||v||^2, alpha = g/||v||, the bf16 copies of the effective kernels, gathered biases)
 * up to date itself across its own calls -- the optimizer rebuilds them for the tensors it writes -- so nothing is re-derived
 * per forward pass; this call marks that state stale and the next eager engine call rebuilds it first.  Not needed after
 * regat_engine_bind.  Before REPLAYING a captured graph after an outside write, call regat_engine_refresh_weights. */
REGAT_API int regat_engine_params_changed(regat_engine* e);
/* Copies the configuration the engine was created with. */
REGAT_API int regat_engine_config(const regat_engine* e, regat_config* cfg);
/* Overrides the 8 sinusoid divisors 1000^(k/8) (default: powf in C).  The Python binding passes the
 * values NumPy computes in fp32 so that stage 1 tracks position_emb.py:98-100 bit for bit.          */
REGAT_API int regat_engine_set_wave_div(regat_engine* e, const float* wave_div_host);
/* Converts the bound grads buffer IN PLACE from dL/dW_eff (what fwd_bwd leaves; the quantity that is
 * all-reduced in data parallel) to the reference's tape.gradient values (dv, dg) of weight_norm.py:41
 * (train.py:111).  Optional -- regat_engine_update accepts either state.                            */
REGAT_API int regat_engine_finalize_grads(regat_engine* e, regat_stream_t stream);
/* Device pointer of a named internal activation (tests and the Python layer mirror): v0, mask, s, Qb,
 * KVb, v1, P, GB, att, pooled, joint, hid, logits, alpha, scal, dv1, ds, dQb, dKVb.                   */
REGAT_API int regat_engine_buffer(const regat_engine* e, const char* name, void** ptr);

/* ------------------------------------------------------------------ DLPack front door ----
 * Same as the two engine calls above, taking DLManagedTensor* (from
 * tf.experimental.dlpack.to_dlpack / torch.utils.dlpack.to_dlpack; capsule "dltensor").
 * Tensors must be kDLCUDA on the current device, fp32, compact row-major.  The capsule stays
 * owned by the caller: the deleter is never called.                                        */
struct DLManagedTensor;
REGAT_API int regat_engine_forward_dl(regat_engine* e, struct DLManagedTensor* features,
                            struct DLManagedTensor* boxes, struct DLManagedTensor* q_att,
                            struct DLManagedTensor* q_last, struct DLManagedTensor* logits_out,
                            regat_stream_t stream);
REGAT_API int regat_engine_train_step_dl(regat_engine* e, struct DLManagedTensor* features,
                               struct DLManagedTensor* boxes, struct DLManagedTensor* q_att,
                               struct DLManagedTensor* q_last, struct DLManagedTensor* target,
                               float lr, int step, float* loss_out, regat_stream_t stream);

/* ------------------------------------------------------------------ question front-end (next row) ---------
 * model/language_model.py:10-174, fp32: what is not a GEMM.  The dense products (x_t W, h U, the two attention FCs and all
 * of their transposes) are regat_gemm calls; tf_vqa_regat_b200/question.py sequences them with the kernels below.
 *
 * regat_q_embed_fwd   language_model.py:33-40 (+ :88-90 when emb2 != NULL, op 'c'): out[r,:] = [emb[tok_r] | emb2[tok_r]],
 *                     zero for tok_r == n_token (the padding index); tables [n_token+1, E]; out [BT, E or 2E].
 * regat_q_embed_bwd   its transpose: demb[tok_r] += dX[r, :E], demb2[tok_r] += dX[r, E:] (atomic; caller zeroes the tables;
 *                     a NULL table is skipped -- the frozen second table, language_model.py:58).
 * regat_q_gru_gates_* one step of keras.layers.GRU (reset_after, blocks z|r|h, language_model.py:106-108): given
 *                     xi = x_t W + b0 (row pitch ld_xi) and hi = h_{t-1} U + b1 [B,3H]:
 *                       z = sig(xz+hz), r = sig(xr+hr), c = tanh(xh + r*hh), h_t = z*h_{t-1} + (1-z)*c
 *                     fwd saves z, r, c and a copy of h_{t-1} ([B,H] each; hp == NULL means the zero initial state);
 *                     bwd takes dh = dh_seq (+ dh_rec) and writes dxi = (dz, dr, da), dhi = (dz, dr, da*r) and
 *                     dhp = dh*z; the caller then accumulates dhi . U^T onto dhp with regat_gemm.
 * regat_q_batch_softmax_*  language_model.py:159-165: P[T,B] = softmax over the BATCH of logits[B,T] transposed.
 * regat_q_pool_*      :165-170: P's memory re-read as w[B,T] (raw reshape); q_att[b,:] = sum_t w[b,t] seq[b,t,:].
 *                     bwd: dseq[b,t,:] = w[b,t] dq_att[b,:] (+ dq_last[b,:] at t = T-1, :120), dW[b,t] = <dq_att[b], seq[b,t]>.
 * regat_q_dot         out[0] += sum a*b (caller zeroes out).   regat_q_wn_alpha: alpha = g / sqrt(max(vv, 1e-12)).
 * regat_q_wn_bwd      weight_norm.py:41 backward from G = dL/dW_eff: dv = alpha (G - Gv v / vv), dg = Gv / sqrt(vv).
 * regat_q_clip_adamax train.py:112-113 for one tensor: clip_by_norm with ||g||^2 = *gsumsq, then Keras Adamax step `step`. */
REGAT_API int regat_q_embed_fwd(const int32_t* tokens, int64_t BT, int n_token, int E, const float* emb, const float* emb2,
                      float* out, regat_stream_t stream);
REGAT_API int regat_q_embed_bwd(const int32_t* tokens, int64_t BT, int n_token, int E, int width, const float* dX, float* demb,
                      float* demb2, regat_stream_t stream);
REGAT_API int regat_q_gru_gates_fwd(int B, int H, const float* xi, int64_t ld_xi, const float* hi, const float* hp, int64_t ld_hp,
                          float* h_out, int64_t ld_h, float* z, float* r, float* c, float* hp_copy, regat_stream_t stream);
REGAT_API int regat_q_gru_gates_bwd(int B, int H, const float* dh_seq, int64_t ld_dseq, const float* dh_rec, const float* z,
                          const float* r, const float* c, const float* hp_copy, const float* hi, float* dxi, int64_t ld_dxi,
                          float* dhi, float* dhp, regat_stream_t stream);
REGAT_API int regat_q_tanh_fwd(float* x, int64_t n, regat_stream_t stream);
REGAT_API int regat_q_tanh_bwd(float* dy, const float* y, int64_t n, regat_stream_t stream);
REGAT_API int regat_q_batch_softmax_fwd(const float* logits, int B, int T, float* P, regat_stream_t stream);
REGAT_API int regat_q_batch_softmax_bwd(const float* P, const float* dP, int B, int T, float* dlogits, regat_stream_t stream);
REGAT_API int regat_q_pool_fwd(const float* w_flat, const float* seq, int B, int T, int H, float* q_att, regat_stream_t stream);
REGAT_API int regat_q_pool_bwd(const float* w_flat, const float* seq, const float* dq_att, const float* dq_last, int B, int T, int H,
                     float* dseq, float* dW, regat_stream_t stream);
REGAT_API int regat_q_dot(const float* a, const float* b, int64_t n, float* out, regat_stream_t stream);
REGAT_API int regat_q_wn_alpha(const float* g, const float* vv, float* alpha, regat_stream_t stream);
REGAT_API int regat_q_wn_bwd(const float* G, const float* v, const float* g, const float* vv, const float* Gv, int64_t n, float* dv,
                   float* dg, regat_stream_t stream);
REGAT_API int regat_q_clip_adamax(float* w, const float* grad, float* m, float* u, int64_t n, const float* gsumsq, float clip,
                        float lr, int step, float beta1, float beta2, float eps, regat_stream_t stream);
/* The embedding tables are read through tf.nn.embedding_lookup (language_model.py:33), so their tape gradient is tf.IndexedSlices and
 * train.py:112-113 treats it as such.  dX [BT, width] is the gradient w.r.t. the (masked) lookup output, the table occupying columns
 * [col0, col0+E); rows of the padding token (tok == n_token) are skipped (their values are zero).
 * regat_q_embed_sumsq       out[0] += sum over occurrences of ||dX[r, col0:col0+E]||^2 -- what tf.clip_by_norm normalises by
 *                           (duplicate tokens NOT summed first); caller zeroes out.
 * regat_q_embed_clip_adamax Keras Adamax, sparse branch, for one table [n_token+1, E]: grad_dense is the scatter-added table
 *                           (regat_q_embed_bwd), clipped by the occurrence norm; m <- b1 m + (1-b1) g; u <- b2 u, then every occurrence
 *                           adds max(u_row, |value|) - u_row of the decayed row (uinc: zeroed scratch of the table's size, left zero);
 *                           table -= lr_t m / (u + eps) over ALL rows. */
REGAT_API int regat_q_embed_sumsq(const int32_t* tokens, int64_t BT, int n_token, int E, int width, int col0, const float* dX,
                        float* out, regat_stream_t stream);
REGAT_API int regat_q_embed_clip_adamax(const int32_t* tokens, int64_t BT, int n_token, int E, int width, int col0, const float* dX,
                              float* table, const float* grad_dense, float* m, float* u, float* uinc, const float* gsumsq,
                              float clip, float lr, int step, float beta1, float beta2, float eps, regat_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* REGAT_H_ */
