// The hot path as one launch sequence on one stream with static buffers:
//   rel_graph_net.py:53-62 (v_relation -> joint_emb -> classifier), train.py:103-113 (loss, gradients,
//   per-tensor clip, Adamax).  No allocation, no host sync, no host-side data dependence, so the
//   whole step can be captured in a CUDA graph by the caller.
//
// Re-associations relative to the reference formulation (all exact in real arithmetic; proven against the
// reference-formulation oracle in tests/):
//   W = g v/||v||_F = alpha*v            -> GEMMs read v (or its bf16 copy), alpha applied in the epilogue
//   [v0 || mask*q] Ws                   -> v0 Ws[:D] + mask (q Ws[D:])          (relation_encoder.py:31-35)
//   grouped conv of (p . s[:M])         -> p . (s[:M] Kc + bc)                  (graph_att_layer.py:110-117)
//   linear((v1 Wva + b) * (q Wqa + b))  -> <v1, Wva ((q Wqa + b) * wl)> + const (fusion.py:47-52, all-linear FCs)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace regat {

struct Layer {
  long long v_off = -1, g_off = -1, b_off = -1;
  int rows = 0, cols = 0;       // kernel viewed as [rows, cols]
  long long lowp_off = -1;      // element offset of the bf16 copy in the workspace
  int lowp_ld = 0;
};

struct Entry { long long off, numel; int layer, kind; };

// A contiguous interval of layers whose gradients become final together in the backward pass: the unit of the gradient exchange
// (data parallel) and of the optimizer, which both run on the `opt` stream behind the rest of the backward pass.
struct OptRange {
  int l_first = 0, l_last = -1;       // layers [l_first, l_last]; empty if l_last < l_first
  long long lo = 0, hi = 0;           // element interval of the flat buffers
  TensorList tl_opt, tl_v, tl_gather; // this range's slices of the engine-wide lists (chunk numbering local to the range)
  int chunks_opt = 0, chunks_v = 0;
  int cbase = 0;                      // first per-chunk slot in `partials`
  int tbase = 0;                      // first per-tensor slot in `stats` / `counters`
  int vbase = 0;                      // first chunk of the range's kernels in the engine-wide ||v||^2 partials (`vpart`)
  int label_local = -1;               // index of the label FC inside tl_v, or -1
};
constexpr int N_RANGES = 4;           // announcement order: BUTD + classifier, attention layers, v2out, self_weights + label FC

int dp_allreduce_f32_impl(const uint64_t* grad_ptrs, uint64_t multicast_ptr, const uint64_t* flag_ptrs, int rank, int world,
                          int64_t offset, int64_t numel, uint32_t epoch, uint32_t* epoch_dev, int blocks, cudaStream_t stream);

}  // namespace regat

using namespace regat;

struct regat_engine {
  regat_config cfg;
  int dtype;
  int max_b, max_n;
  int use_tc;                      // bf16 dense kernels: 1 = tcgen05 (default), 0 = SIMT cross-check (REGAT_GEMM=simt)
  std::vector<Layer> layers;
  std::vector<Entry> entries;      // flat layout order: v, g, [bias] per layer
  long long param_elems = 0;
  int l_v2out = -1, l_self = -1, l_label = -1, l_pos[2], l_q[2], l_k[2], l_out[2];
  int l_va = -1, l_qa = -1, l_lin = -1, l_ve = -1, l_qe = -1, l_c0 = -1, l_c3 = -1;
  TensorList tl_v;  int chunks_v = 0;     // all weight-normed kernels (weight prep)
  TensorList tl_opt; int chunks_opt = 0;  // kernels + biases (optimizer)
  float wave_div[8];
  // bound buffers
  float* params = nullptr; float* grads = nullptr; float* am = nullptr; float* au = nullptr;
  unsigned char* ws = nullptr; long long ws_bytes = 0;
  long long ws_need = 0;
  // workspace carve (byte offsets)
  struct Buf { long long off = -1; };
  Buf lowp, gbias, sumsq, alpha, invn, scal, hyp, stats, partials, vpart, counters, featT, qattT, qlastT, v0, mask, qs, s, strunc, Qb, KVb, v1, P, GB, gate, uqe,
      uw, cb, weff, att, pooled, pv, joint, hid, logits, dlogits, dhid, djoint, dpv, duqe, dpooled, dv1, dweff, dcb, duw,
      dQb, dKVb, ds, dstrunc, dsq, dwc3;
  int a_pad = 0;
  // bf16 mode: layers that share an input sit side by side in one wide bf16 matrix (alpha folded in)
  long long gq_off = 0, gkv_off = 0, guqe_off = 0;      // element offsets of the groups in the lowp buffer
  long long bq_off = 0, bkv_off = 0, buqe_off = 0;      // float offsets of the gathered biases in gbias
  TensorList tl_gather;
  int last_launches = 0;
  int grads_final = 0;
  // Everything derived from the parameters (per-chunk ||v||^2, alpha = g/||v||, the bf16 copies of alpha*v, the gathered biases,
  // the label constant) matches `params`.  This is an invariant of the DEVICE state that every engine call restores before it
  // returns: the optimizer re-derives the state of the tensors it has just written, and a call that finds the flag down (after
  // bind / params_changed) derives everything first.  A captured CUDA graph therefore never depends on a host-side shortcut.
  int weights_ready = 0;
  OptRange rng[N_RANGES];
  // small independent work (BUTD question branch, tiny weight gradients) runs on a side stream, forked / joined with events
  cudaStream_t side = nullptr;
  cudaStream_t opt = nullptr;      // clip + Adamax + re-derived weights, range by range, behind the rest of the backward pass
  cudaStream_t comm = nullptr;     // data parallel: the in-place gradient exchange of each range (runs beside the previous range's optimizer)
  static constexpr int NEV = 24;   // 0-9 forward / backward forks, 10-13 range ready (main), 14 dgrads done, 15 final join, 16-19 range ready (side), 20-23 range exchanged
  cudaEvent_t ev[NEV] = {};
  // data parallel (regat_engine_set_dp): the grads buffer is a symmetric allocation, reduced in place by csrc/dp_exchange.cu
  int dp_world = 1, dp_rank = 0, dp_blocks = 32;
  uint64_t dp_grad_ptrs[16] = {}, dp_flag_ptrs[16] = {}, dp_mc = 0;
  // measurement aid (regat_engine_profile): CUDA events around every dense product of an EAGER call
  struct GemmRec { int M, N, K; cudaEvent_t a, b; };
  std::vector<GemmRec> prof;
  int prof_on = 0;
  regat_grad_ready_fn grad_cb = nullptr;   // data parallel: called when a range of `grads` is final on the stream
  void* grad_cb_user = nullptr;
  template <typename T> T* at(const Buf& b) const { return reinterpret_cast<T*>(ws + b.off); }
  void* atv(const Buf& b) const { return ws + b.off; }
};

namespace regat {
namespace {

constexpr long long ALIGN_ELEMS = 64;   // 256-byte alignment of every tensor in the flat fp32 buffers

long long rup(long long x, long long a) { return (x + a - 1) / a * a; }

// numpy-compatible fp32 1000^(k/8): position_emb.py:98-100 computes np.power in float32.
void fill_wave_div(float* d, int feat_dim) {
  for (int k = 0; k < 8; ++k) d[k] = powf(1000.0f, (8.0f / (float)feat_dim) * (float)k);
}

int add_layer(regat_engine* e, int rows, int cols, bool bias) {
  Layer L;
  L.rows = rows; L.cols = cols;
  long long off = e->param_elems;
  const int li = (int)e->layers.size();
  L.v_off = off; e->entries.push_back({off, (long long)rows * cols, li, 0}); off = rup(off + (long long)rows * cols, ALIGN_ELEMS);
  L.g_off = off; e->entries.push_back({off, 1, li, 1}); off = rup(off + 1, ALIGN_ELEMS);
  if (bias) { L.b_off = off; e->entries.push_back({off, cols, li, 2}); off = rup(off + cols, ALIGN_ELEMS); }
  e->param_elems = off;
  e->layers.push_back(L);
  return li;
}

// Keras-2 variable order of the reference layers; mirrored by tf_vqa_regat_b200/config.py.
void build_layout(regat_engine* e) {
  const regat_config& c = e->cfg;
  const int V = c.v_dim, Q = c.q_dim, D = c.rel_dim, H = c.num_heads, E = c.pos_emb_dim, A = c.num_answers, Hd = c.q_dim;
  if (V != D) e->l_v2out = add_layer(e, V, D, true);                   // relation_encoder.py:52-53
  e->l_self = add_layer(e, D + Q, D, true);                            // graph_att_net.py:24
  e->l_label = add_layer(e, 1, 1, c.label_bias != 0);                  // graph_att_net.py:25
  for (int d = 0; d < c.dir_num; ++d) {                                // graph_att_layer.py:25-37
    e->l_pos[d] = add_layer(e, E, H, true);
    e->l_q[d] = add_layer(e, D, D, true);
    e->l_k[d] = add_layer(e, D, D, true);
    e->l_out[d] = add_layer(e, D, D, true);                            // Conv2D kernel [1,1,D,D]
  }
  e->l_va = add_layer(e, D, Hd, true);                                 // fusion.py:15-20
  e->l_qa = add_layer(e, Q, Hd, true);
  e->l_lin = add_layer(e, Hd, 1, true);
  e->l_ve = add_layer(e, D, Hd, true);
  e->l_qe = add_layer(e, Q, Hd, true);
  e->l_c0 = add_layer(e, Hd, 2 * Hd, true);                            // classifier.py:14-19
  e->l_c3 = add_layer(e, 2 * Hd, A, true);
}

long long carve(regat_engine* e) {
  const regat_config& c = e->cfg;
  const long long B = e->max_b, N = e->max_n, M = std::min<long long>(c.nongt_dim, N);
  const long long R = B * N, Rm = B * M;
  const long long V = c.v_dim, Q = c.q_dim, D = c.rel_dim, H = c.num_heads, A = c.num_answers, Hd = c.q_dim, dirs = c.dir_num;
  const long long es = dtype_size(e->dtype);
  long long off = 0;
  auto take = [&](regat_engine::Buf& b, long long bytes) { b.off = off; off = rup(off + bytes, 256); };
  e->a_pad = (int)rup(A, 64);
  // bf16 copies of the effective kernels alpha*v (bf16 mode only); row pitch padded to a multiple of 8 elements.
  // Groups: [Q_0|Q_1] (ld dirs*D), [K_0|K_1|V'_0|V'_1] (ld 2*dirs*D), [q2attention|question_embed] (ld 2*Hd).
  long long lowp_elems = 0;
  auto place = [&](int l, long long off, int ld) { e->layers[l].lowp_off = off; e->layers[l].lowp_ld = ld; };
  e->gq_off = lowp_elems;
  for (int d = 0; d < dirs; ++d) place(e->l_q[d], e->gq_off + d * D, (int)(dirs * D));
  lowp_elems = rup(lowp_elems + D * dirs * D, 128);
  e->gkv_off = lowp_elems;
  for (int d = 0; d < dirs; ++d) {
    place(e->l_k[d], e->gkv_off + d * D, (int)(2 * dirs * D));
    place(e->l_out[d], e->gkv_off + (dirs + d) * D, (int)(2 * dirs * D));
  }
  lowp_elems = rup(lowp_elems + D * 2 * dirs * D, 128);
  e->guqe_off = lowp_elems;
  place(e->l_qa, e->guqe_off, (int)(2 * Hd));
  place(e->l_qe, e->guqe_off + Hd, (int)(2 * Hd));
  lowp_elems = rup(lowp_elems + Q * 2 * Hd, 128);
  for (auto& L : e->layers) {
    if (L.lowp_off >= 0) continue;
    L.lowp_ld = (int)rup(L.cols, 8);
    L.lowp_off = lowp_elems;
    lowp_elems = rup(lowp_elems + (long long)L.rows * L.lowp_ld, 128);
  }
  e->bq_off = 0; e->bkv_off = dirs * D; e->buqe_off = 3 * dirs * D;
  take(e->gbias, (3 * dirs * D + 2 * Hd) * 4);
  take(e->lowp, e->dtype == REGAT_BF16 ? lowp_elems * 2 : 0);
  const long long nl = (long long)e->layers.size();
  take(e->sumsq, nl * 4); take(e->alpha, nl * 4); take(e->invn, nl * 4);
  take(e->scal, 64 * 4);                       // [0]=label const c, [1]=dc, [2]=loss, [3]=score
  take(e->hyp, 64 * 4);                        // Hyper {lr, step, lr_t} at [0], the exchange's call counter (uint32) at [8]
  take(e->stats, 2 * MAX_TENSORS * 4);
  take(e->partials, 2 * 8192 * 4);             // per-chunk partial sums of the optimizer reductions
  take(e->counters, MAX_TENSORS * 4);          // per-tensor "chunks finished" counters of the optimizer reduction (self-resetting)
  take(e->vpart, 8192 * 4);                    // per-chunk ||v||^2 (written by wn_prepare or, after a step, by the update itself)
  take(e->featT, e->dtype == REGAT_BF16 ? R * V * 2 : 0);
  take(e->qattT, e->dtype == REGAT_BF16 ? B * Q * 2 : 0);
  take(e->qlastT, e->dtype == REGAT_BF16 ? B * Q * 2 : 0);
  take(e->v0, e->l_v2out >= 0 ? R * D * es : 0);
  take(e->mask, R * 4);
  take(e->qs, B * D * 4);
  take(e->s, R * D * es); take(e->strunc, Rm * D * es);
  take(e->Qb, R * dirs * D * es); take(e->KVb, Rm * 2 * dirs * D * es);
  take(e->v1, R * D * es);
  take(e->P, B * dirs * H * N * M * 4); take(e->GB, B * dirs * H * N * M * 4);
  take(e->gate, R * H * 8);
  take(e->uqe, B * 2 * Hd * es); take(e->uw, B * Hd * es); take(e->cb, B * 4);
  take(e->weff, B * D * es); take(e->att, B * N * 4); take(e->pooled, B * D * es);
  take(e->pv, B * Hd * es); take(e->joint, B * Hd * es); take(e->hid, B * 2 * Hd * es);
  take(e->logits, B * e->a_pad * 4);
  take(e->dwc3, e->dtype == REGAT_BF16 ? 2 * Hd * (long long)e->a_pad * 4 : 0);
  take(e->dlogits, B * e->a_pad * es);
  take(e->dhid, B * 2 * Hd * es); take(e->djoint, B * Hd * es); take(e->dpv, B * Hd * es); take(e->duqe, B * 2 * Hd * es);
  take(e->dpooled, B * D * es); take(e->dv1, R * D * es); take(e->dweff, B * D * es); take(e->dcb, B * 4);
  take(e->duw, B * Hd * es);
  take(e->dQb, R * dirs * D * es); take(e->dKVb, Rm * 2 * dirs * D * es);
  take(e->ds, R * D * es); take(e->dstrunc, Rm * D * es); take(e->dsq, B * D * es);
  return off;
}

void build_lists(regat_engine* e) {
  TensorList& tv = e->tl_v;
  memset(&tv, 0, sizeof(tv));
  TensorList& to = e->tl_opt;
  memset(&to, 0, sizeof(to));
  for (size_t l = 0; l < e->layers.size(); ++l) {
    const Layer& L = e->layers[l];
    int i = tv.n++;
    tv.off[i] = L.v_off; tv.numel[i] = (long long)L.rows * L.cols; tv.g_off[i] = L.g_off; tv.cols[i] = L.cols;
    tv.off_lowp[i] = L.lowp_off; tv.ld_lowp[i] = L.lowp_ld; tv.kind[i] = 0; tv.layer[i] = (int)l;
    int j = to.n++;
    to.off[j] = L.v_off; to.numel[j] = (long long)L.rows * L.cols; to.g_off[j] = L.g_off; to.cols[j] = L.cols; to.kind[j] = 0;
    to.layer[j] = (int)l;
    if (L.b_off >= 0) {
      j = to.n++;
      to.off[j] = L.b_off; to.numel[j] = L.cols; to.g_off[j] = -1; to.cols[j] = L.cols; to.kind[j] = 1; to.layer[j] = (int)l;
    }
  }
  e->chunks_v = build_tensor_list(tv);
  e->chunks_opt = build_tensor_list(to);
  for (int j = 0, i = 0; j < to.n; ++j) {       // kernels appear in the same order in both lists
    to.vchunk_start[j] = 0;
    if (to.kind[j] == 0) to.vchunk_start[j] = tv.chunk_start[i++];
  }
  TensorList& tg = e->tl_gather;
  memset(&tg, 0, sizeof(tg));
  const long long D = e->cfg.rel_dim, Hd = e->cfg.q_dim;
  const int dirs = e->cfg.dir_num;
  auto add = [&](int l, long long dst) {
    const Layer& L = e->layers[l];
    if (L.b_off < 0) return;
    tg.off[tg.n] = L.b_off; tg.numel[tg.n] = L.cols; tg.off_lowp[tg.n] = dst; tg.layer[tg.n] = l; ++tg.n;
  };
  for (int d = 0; d < dirs; ++d) {
    add(e->l_q[d], e->bq_off + d * D);
    add(e->l_k[d], e->bkv_off + d * D);
    add(e->l_out[d], e->bkv_off + (dirs + d) * D);
  }
  add(e->l_qa, e->buqe_off);
  add(e->l_qe, e->buqe_off + Hd);

  // ---- ranges, in the order the backward pass finishes them
  const int nl = (int)e->layers.size();
  e->rng[0].l_first = e->l_va; e->rng[0].l_last = e->l_c3;
  e->rng[1].l_first = e->l_pos[0]; e->rng[1].l_last = e->l_out[dirs - 1];
  if (e->l_v2out >= 0) { e->rng[2].l_first = e->l_v2out; e->rng[2].l_last = e->l_v2out; }
  e->rng[3].l_first = e->l_self; e->rng[3].l_last = e->l_label;
  int cbase = 0, tbase = 0;
  for (int r = 0; r < N_RANGES; ++r) {
    OptRange& R = e->rng[r];
    memset(&R.tl_opt, 0, sizeof(TensorList)); memset(&R.tl_v, 0, sizeof(TensorList)); memset(&R.tl_gather, 0, sizeof(TensorList));
    if (R.l_last < R.l_first) continue;
    R.lo = e->layers[R.l_first].v_off;
    R.hi = R.l_last + 1 < nl ? e->layers[R.l_last + 1].v_off : e->param_elems;
    auto copy = [](TensorList& d, const TensorList& src, int j) {
      const int i = d.n++;
      d.off[i] = src.off[j]; d.numel[i] = src.numel[j]; d.g_off[i] = src.g_off[j]; d.off_lowp[i] = src.off_lowp[j];
      d.ld_lowp[i] = src.ld_lowp[j]; d.cols[i] = src.cols[j]; d.kind[i] = src.kind[j]; d.layer[i] = src.layer[j];
      return i;
    };
    for (int j = 0; j < to.n; ++j)
      if (to.layer[j] >= R.l_first && to.layer[j] <= R.l_last) {
        const int i = copy(R.tl_opt, to, j);
        R.tl_opt.vchunk_start[i] = to.vchunk_start[j];        // engine-wide ||v||^2 chunk numbering
      }
    for (int j = 0; j < tv.n; ++j)
      if (tv.layer[j] >= R.l_first && tv.layer[j] <= R.l_last) {
        const int i = copy(R.tl_v, tv, j);
        if (tv.layer[j] == e->l_label) R.label_local = i;
      }
    for (int j = 0; j < tg.n; ++j)
      if (tg.layer[j] >= R.l_first && tg.layer[j] <= R.l_last) copy(R.tl_gather, tg, j);
    R.chunks_opt = build_tensor_list(R.tl_opt);
    R.chunks_v = build_tensor_list(R.tl_v);
    R.vbase = tv.chunk_start[R.l_first];                       // tl_v holds one entry per layer, in layer order
    R.cbase = cbase; R.tbase = tbase;
    cbase += R.chunks_opt; tbase += R.tl_opt.n;
  }
}

// zeroes the atomically accumulated gradient slots; with `hyp` it also opens the optimizer step (++step, lr_t) -- one launch
__global__ void zero_list_kernel(float* buf, TensorList tl, Hyper* hyp, float beta1) {
  for (int l = blockIdx.x; l < tl.n; l += gridDim.x)
    for (long long i = threadIdx.x; i < tl.numel[l]; i += blockDim.x) buf[tl.off[l] + i] = 0.f;
  if (hyp && blockIdx.x == 0 && threadIdx.x == 0) hyper_tick(hyp, beta1);
}

struct Ctx {
  regat_engine* e;
  cudaStream_t st;
  int B, N, M, R, Rm;
  const float* features; const float* boxes; const float* q_att; const float* q_last;
  bool fused_opt = false;    // train step: exchange + optimizer of each gradient range on the `opt` stream as soon as it is final
  bool chain = false;        // the batch-sized tail (pv .. logits, loss, dhid .. dpooled) runs as ONE chained launch in backward()
  float* logits_out = nullptr;   // fwd_bwd caller's logits pointer, served after the chain
};

EpiArgs epi0() { EpiArgs x; memset(&x, 0, sizeof(x)); return x; }

// brackets one dense product with timing events when regat_engine_profile is on (eager calls only)
struct ProfScope {
  regat_engine* e; cudaStream_t st; int idx = -1;
  ProfScope(regat_engine* e_, cudaStream_t st_, int M, int N, int K) : e(e_), st(st_) {
    if (!e->prof_on) return;
    regat_engine::GemmRec r{M, N, K, nullptr, nullptr};
    if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
    cudaEventRecord(r.a, st);
    idx = (int)e->prof.size();
    e->prof.push_back(r);
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(e->prof[idx].b, st); }
};

int dense(regat_engine* e, cudaStream_t st, bool tA, bool tB, int M, int N, int K, const void* A, int lda, const void* Bm,
          int ldb, void* C, int ldc, int c_dtype, const EpiArgs& ep, int split_k = 1) {
  ProfScope prof(e, st, M, N, K);
  if (e->dtype == REGAT_F32) return gemm_simt(REGAT_F32, tA, tB, M, N, K, A, lda, Bm, ldb, C, ldc, REGAT_F32, ep, st);
  if (e->use_tc && gemm_tc_supported(tA, tB, M, N, K, A, lda, Bm, ldb)) {
    // products on the side stream are off the dependency chain: they may not take every SM from the main stream's kernels
    // (a persistent launch holds its SMs' shared memory until it ends).  REGAT_SIDE_CAP = CTAs they may use, 0 = no limit.
    static const int side_cap = [] { const char* s = getenv("REGAT_SIDE_CAP"); return s ? atoi(s) : 0; }();
    return gemm_tc(tA, tB, M, N, K, A, lda, Bm, ldb, C, ldc, c_dtype, ep, split_k, st, 0, nullptr, st == e->side ? side_cap : 0);
  }
  return gemm_simt(REGAT_BF16, tA, tB, M, N, K, A, lda, Bm, ldb, C, ldc, c_dtype, ep, st);
}
// weight gradients of side-by-side layers: one product whose column blocks land in the layers' own gradient slots
int dense_scatter(regat_engine* e, cudaStream_t st, int M, int N, int K, const void* A, int lda, const void* Bm, int ldb, float* C,
                  int ldc, int block_cols, const long long* block_off) {
  ProfScope prof(e, st, M, N, K);
  return gemm_tc(1, 0, M, N, K, A, lda, Bm, ldb, C, ldc, REGAT_F32, epi0(), 1, st, block_cols, block_off);
}

// kernel operand of layer l: fp32 master weights in parity mode, bf16 copy otherwise
const void* W(const regat_engine* e, int l, long long row = 0) {
  const Layer& L = e->layers[l];
  if (e->dtype == REGAT_F32) return e->params + L.v_off + row * L.cols;
  return reinterpret_cast<const bf16*>(e->ws + e->lowp.off) + L.lowp_off + row * L.lowp_ld;
}
int ldW(const regat_engine* e, int l) { return e->dtype == REGAT_F32 ? e->layers[l].cols : e->layers[l].lowp_ld; }
const float* biasp(const regat_engine* e, int l) { return e->layers[l].b_off >= 0 ? e->params + e->layers[l].b_off : nullptr; }
float* gradW(const regat_engine* e, int l, long long row = 0) { return e->grads + e->layers[l].v_off + row * e->layers[l].cols; }
float* gradB(const regat_engine* e, int l) { return e->layers[l].b_off >= 0 ? e->grads + e->layers[l].b_off : nullptr; }
const float* alphap(const regat_engine* e, int l) { return e->at<float>(e->alpha) + l; }
// alpha for GEMM epilogues: folded into the bf16 weight copies, explicit in fp32 parity mode
const float* alpha_epi(const regat_engine* e, int l) { return e->dtype == REGAT_F32 ? alphap(e, l) : nullptr; }
const bf16* lowp_at(const regat_engine* e, long long off) { return reinterpret_cast<const bf16*>(e->ws + e->lowp.off) + off; }

// y = act(alpha*(x W) + b)
int fc_fwd(regat_engine* e, cudaStream_t st, int l, long long w_row0, int rows, int K, const void* x, int ldx, void* y, int ldy,
           int y_dtype, bool relu, bool with_bias = true) {
  EpiArgs ep = epi0();
  ep.alpha = alpha_epi(e, l);
  ep.bias = with_bias ? biasp(e, l) : nullptr;
  ep.relu = relu;
  return dense(e, st, false, false, rows, e->layers[l].cols, K, x, ldx, W(e, l, w_row0), ldW(e, l), y, ldy, y_dtype, ep);
}
// dx (+)= alpha * (dy W^T)   [optionally gated]
int fc_dgrad(regat_engine* e, cudaStream_t st, int l, long long w_row0, int rows, int K_in, const void* dy, int lddy, void* dx,
             int lddx, int dx_dtype, bool accumulate, const void* gate = nullptr, int gate_ld = 0) {
  EpiArgs ep = epi0();
  ep.alpha = alpha_epi(e, l);
  ep.accumulate = accumulate;
  ep.gate = gate; ep.gate_ld = gate_ld;
  return dense(e, st, false, true, rows, K_in, e->layers[l].cols, dy, lddy, W(e, l, w_row0), ldW(e, l), dx, lddx, dx_dtype, ep);
}
// dW_eff[w_row0 : w_row0+K_in, :] = x^T dy  (fp32, into the grads buffer);  db += colsum(dy)
// the bias gradient (a column sum of dy) is either launched right away or queued in `batch` for one multi-problem launch
int bias_grad(regat_engine* e, cudaStream_t st, const void* dy, int lddy, int rows, int cols, float* out, ColsumBatch* batch) {
  if (!out) return REGAT_OK;
  if (batch && aligned16(dy) && lddy % 8 == 0 && cols % 8 == 0 && batch->n < 12) { batch->add(dy, lddy, rows, cols, out); return REGAT_OK; }
  return k_colsum(e->dtype, dy, lddy, rows, cols, out, st);
}
int fc_wgrad(regat_engine* e, cudaStream_t st, int l, long long w_row0, int rows, int K_in, const void* x, int ldx, const void* dy,
             int lddy, bool with_bias, ColsumBatch* batch = nullptr) {
  EpiArgs ep = epi0();
  REGAT_TRY(dense(e, st, true, false, K_in, e->layers[l].cols, rows, x, ldx, dy, lddy, gradW(e, l, w_row0), e->layers[l].cols,
                  REGAT_F32, ep));
  if (with_bias) REGAT_TRY(bias_grad(e, st, dy, lddy, rows, e->layers[l].cols, gradB(e, l), batch));
  return REGAT_OK;
}

// alpha = g/||v|| (from the per-chunk ||v||^2 partials), gathered biases, label constant and -- bf16 mode -- the bf16 copies of
// alpha*v, for the layers of one range (r < 0: all layers)
int derive_weights(regat_engine* e, int r, cudaStream_t st) {
  const bool all = r < 0;
  const TensorList& tv = all ? e->tl_v : e->rng[r].tl_v;
  const TensorList& tg = all ? e->tl_gather : e->rng[r].tl_gather;
  const int l0 = all ? 0 : e->rng[r].l_first;
  const int chunks = all ? e->chunks_v : e->rng[r].chunks_v;
  const int vbase = all ? 0 : e->rng[r].vbase;
  const int label_local = all ? e->l_label : e->rng[r].label_local;
  if (tv.n == 0) return REGAT_OK;
  // the same one-block kernel also gathers the biases of side-by-side layers and evaluates the label-FC constant
  // (graph_att_net.py:71)
  const Layer& LL = e->layers[e->l_label];
  REGAT_TRY(k_wn_alpha(e->params, tv, e->at<float>(e->sumsq) + l0, e->at<float>(e->alpha) + l0, e->at<float>(e->invn) + l0, st,
                       e->at<float>(e->vpart) + vbase, (e->dtype == REGAT_BF16 && tg.n) ? &tg : nullptr, e->at<float>(e->gbias),
                       label_local, LL.v_off, LL.b_off, label_local >= 0 ? e->at<float>(e->scal) : nullptr));
  static const int dbg_nocopy = [] { const char* s = getenv("REGAT_OPT_DEBUG"); return (s && strstr(s, "nocopy")) ? 1 : 0; }();
  if (e->dtype == REGAT_BF16 && !(dbg_nocopy && !all)) REGAT_TRY(k_wn_scaled_copy(e->params, tv, chunks, e->at<float>(e->alpha), e->atv(e->lowp), st));
  return REGAT_OK;
}

// Full derivation from the parameters (first call after bind / params_changed).
int prepare_weights(regat_engine* e, cudaStream_t st) {
  if (e->weights_ready) return REGAT_OK;
  REGAT_REQUIRE(e->chunks_v <= 8192 && e->chunks_opt <= 8192, REGAT_ERR_UNSUPPORTED, "engine: parameter buffer too large for the partial-sum scratch");
  REGAT_TRY(k_wn_prepare(e->params, e->tl_v, e->chunks_v, e->at<float>(e->sumsq), nullptr, st, e->at<float>(e->vpart)));
  REGAT_TRY(derive_weights(e, -1, st));
  e->weights_ready = 1;
  return REGAT_OK;
}

OptHyper opt_hyper(const regat_engine* e) {
  OptHyper hp;
  hp.lr_t = 0.f; hp.beta1 = e->cfg.beta1; hp.beta2 = e->cfg.beta2; hp.eps = e->cfg.eps; hp.clip = e->cfg.grad_clip;
  hp.grads_are_final = e->grads_final;
  hp.lr_t_dev = &e->at<Hyper>(e->hyp)->lr_t;
  return hp;
}

// clip + Adamax of one range (r < 0: everything), then the derived state of the tensors just written
// (Measured and dropped: capping the grids of these launches so that they stream at a fraction of the HBM bandwidth beside the
// backward pass.  At 296 / 148 / 74 CTAs the step went from 0.98 ms to 1.15 / 1.45 / 2.02 ms: the optimizer has less slack than
// its traffic needs, it has to run at full width.)
int optimize_range(regat_engine* e, int r, cudaStream_t st) {
  const bool all = r < 0;
  const TensorList& to = all ? e->tl_opt : e->rng[r].tl_opt;
  const int chunks = all ? e->chunks_opt : e->rng[r].chunks_opt;
  const int cbase = all ? 0 : e->rng[r].cbase, tbase = all ? 0 : e->rng[r].tbase;
  if (to.n == 0) return REGAT_OK;
  float* stats = e->at<float>(e->stats) + 2 * tbase;
  // (Measured and dropped: one launch with grid-wide barriers for the small ranges at the tail of the step -- statistics, update,
  // alpha and bf16 copies bit-identical to the four kernels -- made the step 5 us SLOWER: its blocks cannot become resident beside
  // the weight-gradient GEMMs that are still running, while the four small grids slip into the gaps.)
  static const int dbg_noreduce = [] { const char* s = getenv("REGAT_OPT_DEBUG"); return (s && strstr(s, "noreduce")) ? 1 : 0; }();
  if (!(dbg_noreduce && !all))
  REGAT_TRY(k_opt_reduce(e->params, e->grads, to, chunks, e->at<float>(e->partials) + 2 * cbase, stats, st,
                         e->at<unsigned int>(e->counters) + tbase));
  // the update leaves ||v_new||^2 per chunk in `vpart` (engine-wide chunk numbering)
  REGAT_TRY(k_opt_update(e->params, e->grads, e->am, e->au, to, chunks, stats, e->at<float>(e->alpha), e->at<float>(e->invn),
                         opt_hyper(e), st, e->at<float>(e->vpart)));
  return derive_weights(e, r, st);
}

// bf16 engine, M <= 24 keys: packed-bf16 fast path of the fused attention kernels (P buffer = probabilities -> dz, GB buffer = rz)
bool attn_fast(const regat_engine* e, int N) { return e->dtype == REGAT_BF16 && regat_geoattn_fast_supported(N, e->cfg.nongt_dim) != 0; }

// The side stream (bias sums, geometry reduction, off-chain products) runs at the lowest priority; the optimizer and exchange
// streams take the priority of the caller's stream.  A caller that runs the step on a high-priority stream (bench.py,
// GraphedTrainStep) thereby ranks the dependency chain and the optimizer above the side work when SM slots free up
// (REGAT_OPT_PRIO=low: optimizer at the lowest priority too).
int ensure_side(regat_engine* e, cudaStream_t caller) {
  if (e->side) return REGAT_OK;
  int lo = 0, hi = 0, pr = 0;
  REGAT_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  if (caller == nullptr || cudaStreamGetPriority(caller, &pr) != cudaSuccess) { pr = lo; (void)cudaGetLastError(); }
  const char* op = getenv("REGAT_OPT_PRIO");
  int opt_pr = (op && strcmp(op, "low") == 0) ? lo : pr;
  if (op && strcmp(op, "high") == 0) opt_pr = hi;      // numerically smallest = highest
  REGAT_CUDA(cudaStreamCreateWithPriority(&e->side, cudaStreamNonBlocking, lo));
  REGAT_CUDA(cudaStreamCreateWithPriority(&e->opt, cudaStreamNonBlocking, opt_pr));
  REGAT_CUDA(cudaStreamCreateWithPriority(&e->comm, cudaStreamNonBlocking, opt_pr));
  for (int i = 0; i < regat_engine::NEV; ++i) REGAT_CUDA(cudaEventCreateWithFlags(&e->ev[i], cudaEventDisableTiming));
  return REGAT_OK;
}
// side stream waits for everything enqueued on `from` so far
int fork_to(cudaStream_t from, cudaStream_t to, cudaEvent_t ev) {
  REGAT_CUDA(cudaEventRecord(ev, from));
  REGAT_CUDA(cudaStreamWaitEvent(to, ev, 0));
  return REGAT_OK;
}

// bias / pair_pos_fc / label gradients, the loss and the score are accumulated with atomics: zero those slots (the kernels'
// gradients are overwritten).  With `tick` the same launch opens the optimizer step (++step, lr_t).  Issued on the side stream at
// the start of a training forward pass, off the critical path; every consumer is ordered behind it (same stream, or the join
// before the pooling kernel).
int zero_accumulators(regat_engine* e, cudaStream_t st, bool tick) {
  const int dirs = e->cfg.dir_num;
  TensorList z;
  memset(&z, 0, sizeof(z));
  for (size_t l = 0; l < e->layers.size(); ++l) {
    const Layer& L = e->layers[l];
    if (L.b_off >= 0) { z.off[z.n] = L.b_off; z.numel[z.n] = L.cols; ++z.n; }
    z.off[z.n] = L.g_off; z.numel[z.n] = 1; ++z.n;
  }
  {
    const Layer& L = e->layers[e->l_lin];
    z.off[z.n] = L.v_off; z.numel[z.n] = (long long)L.rows * L.cols; ++z.n;
  }
  for (int d = 0; d < dirs; ++d) {
    const Layer& L = e->layers[e->l_pos[d]];
    z.off[z.n] = L.v_off; z.numel[z.n] = (long long)L.rows * L.cols; ++z.n;
  }
  zero_list_kernel<<<z.n, 128, 0, st>>>(e->grads, z, tick ? e->at<Hyper>(e->hyp) : nullptr, e->cfg.beta1);
  REGAT_POST_LAUNCH();
  REGAT_CUDA(cudaMemsetAsync(e->at<float>(e->scal) + 1, 0, 3 * sizeof(float), st));   // dc, loss, score
  return REGAT_OK;
}

int forward(Ctx& c, bool training, float* logits_out, float* att_out) {
  regat_engine* e = c.e;
  cudaStream_t st = c.st;
  const regat_config& cf = e->cfg;
  const int dt = e->dtype, B = c.B, N = c.N, M = c.M, R = c.R, Rm = c.Rm;
  const int V = cf.v_dim, Q = cf.q_dim, D = cf.rel_dim, H = cf.num_heads, A = cf.num_answers, Hd = cf.q_dim, dirs = cf.dir_num;

  REGAT_TRY(ensure_side(e, st));
  cudaStream_t sd = e->side;
  const size_t es = dtype_size(dt);
  // activations in the compute dtype: the casts do not depend on the weights, so they run on the side stream while the main
  // stream derives alpha and the bf16 kernels from the parameters (both HBM-bound, neither saturates the memory system alone)
  const void* feat = c.features; const void* qatt = c.q_att; const void* qlast = c.q_last;
  if (dt == REGAT_BF16) {
    // the feature cast is the first thing on the step's critical path (v2out waits for it): it goes first on the side stream and
    // the main stream waits for it alone; the question casts and the accumulator zeroing, which only later side-stream work and
    // the backward pass need, queue behind it
    REGAT_TRY(fork_to(st, sd, e->ev[2]));
    REGAT_TRY(k_cast(REGAT_BF16, c.features, e->atv(e->featT), (long long)R * V, sd));
    REGAT_CUDA(cudaEventRecord(e->ev[3], sd));                       // bf16 features ready
    REGAT_TRY(k_cast(REGAT_BF16, c.q_last, e->atv(e->qlastT), (long long)B * Q, sd));
    REGAT_TRY(k_cast(REGAT_BF16, c.q_att, e->atv(e->qattT), (long long)B * Q, sd));
    feat = e->atv(e->featT); qatt = e->atv(e->qattT); qlast = e->atv(e->qlastT);
  }
  if (training && e->grads) {
    REGAT_TRY(fork_to(st, sd, e->ev[2]));
    REGAT_TRY(zero_accumulators(e, sd, c.fused_opt));
  }
  REGAT_TRY(prepare_weights(e, st));
  // ---- side stream: the question branches depend only on q_last / q_att and the weights (fusion.py:37,47-52)
  REGAT_TRY(fork_to(st, sd, e->ev[0]));
  if (dt == REGAT_BF16) REGAT_CUDA(cudaStreamWaitEvent(st, e->ev[3], 0));      // main stream: the bf16 features are ready
  // alpha of the pair_pos_fc layers must be contiguous per direction for the attention kernels: gathered here, off the main
  // stream's chain of encoder products (the main stream joins at ev[4], before the s product)
  for (int d = 0; d < dirs; ++d)
    REGAT_CUDA(cudaMemcpyAsync(e->at<float>(e->scal) + 8 + d, alphap(e, e->l_pos[d]), sizeof(float), cudaMemcpyDeviceToDevice, sd));
  {  // qs = q_att Ws[D:]  (the question half of self_weights' input, relation_encoder.py:31-35); consumed by the s GEMM below
    EpiArgs ep = epi0();
    REGAT_TRY(dense(e, sd, false, false, B, D, Q, qatt, Q, W(e, e->l_self, D), ldW(e, e->l_self), e->atv(e->qs), D, REGAT_F32, ep));
    REGAT_TRY(fork_to(sd, sd, e->ev[4]));                            // record: qs ready
  }
  if (dt == REGAT_BF16 && e->use_tc) {
    EpiArgs ep = epi0();
    ep.bias = e->at<float>(e->gbias) + e->buqe_off;
    REGAT_TRY(dense(e, sd, false, false, B, 2 * Hd, Q, qlast, Q, lowp_at(e, e->guqe_off), 2 * Hd, e->atv(e->uqe), 2 * Hd, dt, ep));
  } else {
    REGAT_TRY(fc_fwd(e, sd, e->l_qa, 0, B, Q, qlast, Q, e->atv(e->uqe), 2 * Hd, dt, false));
    REGAT_TRY(fc_fwd(e, sd, e->l_qe, 0, B, Q, qlast, Q, e->at<unsigned char>(e->uqe) + (size_t)Hd * es, 2 * Hd, dt, false));
  }
  REGAT_TRY(k_butd_prep(dt, e->atv(e->uqe), 2 * Hd, e->params + e->layers[e->l_lin].v_off, alphap(e, e->l_lin), biasp(e, e->l_va),
                        biasp(e, e->l_lin), e->atv(e->uw), e->at<float>(e->cb), B, Hd, sd));
  {  // weff = alpha_va * (uw Wva^T)
    EpiArgs ep = epi0();
    ep.alpha = alpha_epi(e, e->l_va);
    REGAT_TRY(dense(e, sd, false, true, B, D, Hd, e->atv(e->uw), Hd, W(e, e->l_va), ldW(e, e->l_va), e->atv(e->weff), D, dt, ep));
  }
  // ---- main stream: the encoder
  // v0 = relu(v2out(visual))                                            relation_encoder.py:78-79
  const void* v0 = feat;
  if (e->l_v2out >= 0) {
    REGAT_TRY(fc_fwd(e, st, e->l_v2out, 0, R, V, feat, V, e->atv(e->v0), D, dt, true));
    v0 = e->atv(e->v0);
  }
  // mask = (sum_d v0 != 0);  s = self_weights([v0 || mask*q])            relation_encoder.py:13-37, graph_att_net.py:58
  REGAT_TRY(k_rowmask(dt, v0, R, D, e->at<float>(e->mask), st));
  {
    EpiArgs ep = epi0();
    REGAT_CUDA(cudaStreamWaitEvent(st, e->ev[4], 0));                // qs from the side stream
    ep.alpha = alpha_epi(e, e->l_self); ep.bias = biasp(e, e->l_self);
    ep.addend = e->at<float>(e->qs); ep.addend_ld = D; ep.addend_rows = N; ep.row_scale = e->at<float>(e->mask);
    ep.c2 = e->atv(e->strunc); ep.c2_ld = D; ep.c2_rows_in = N; ep.c2_rows_keep = M;
    REGAT_TRY(dense(e, st, false, false, R, D, D, v0, D, W(e, e->l_self, 0), ldW(e, e->l_self), e->atv(e->s), D, dt, ep));
  }
  // per direction: Q = query(s), K = key(s[:, :M]), V' = s[:, :M] Kc + bc   graph_att_layer.py:47,55,112-117
  if (dt == REGAT_BF16 && e->use_tc) {
    // one wide GEMM per input: [Q_0|Q_1] = s [W_q0|W_q1] + b,  [K_0|K_1|V'_0|V'_1] = s[:, :M] [W_k0|W_k1|Kc_0|Kc_1] + b
    EpiArgs ep = epi0();
    ep.bias = e->at<float>(e->gbias) + e->bq_off;
    REGAT_TRY(dense(e, st, false, false, R, dirs * D, D, e->atv(e->s), D, lowp_at(e, e->gq_off), dirs * D, e->atv(e->Qb), dirs * D, dt, ep));
    ep.bias = e->at<float>(e->gbias) + e->bkv_off;
    REGAT_TRY(dense(e, st, false, false, Rm, 2 * dirs * D, D, e->atv(e->strunc), D, lowp_at(e, e->gkv_off), 2 * dirs * D, e->atv(e->KVb),
                    2 * dirs * D, dt, ep));
  } else {
    for (int d = 0; d < dirs; ++d) {
      REGAT_TRY(fc_fwd(e, st, e->l_q[d], 0, R, D, e->atv(e->s), D, e->at<unsigned char>(e->Qb) + (size_t)d * D * es, dirs * D, dt, false));
      REGAT_TRY(fc_fwd(e, st, e->l_k[d], 0, Rm, D, e->atv(e->strunc), D, e->at<unsigned char>(e->KVb) + (size_t)d * D * es, 2 * dirs * D, dt, false));
      REGAT_TRY(fc_fwd(e, st, e->l_out[d], 0, Rm, D, e->atv(e->strunc), D, e->at<unsigned char>(e->KVb) + (size_t)(dirs + d) * D * es, 2 * dirs * D, dt, false));
    }
  }
  // fused geometry-bias attention + relu + residual -> v1
  {
    const Layer& P0 = e->layers[e->l_pos[0]];
    const long long wstride = dirs > 1 ? e->layers[e->l_pos[1]].v_off - P0.v_off : 0;
    const long long bstride = dirs > 1 ? e->layers[e->l_pos[1]].b_off - P0.b_off : 0;
    const float* ag = e->at<float>(e->scal) + 8;                     // gathered on the side stream above
    if (attn_fast(e, N))
      REGAT_TRY(regat_geoattn_fwd_fast(B, N, cf.nongt_dim, D, H, dirs, cf.pos_emb_dim, e->atv(e->Qb), e->atv(e->KVb), c.boxes, e->wave_div,
                                       e->params + P0.v_off, wstride, ag, e->params + P0.b_off, bstride, e->at<float>(e->scal),
                                       e->atv(e->s), v0, cf.residual, e->atv(e->v1), training ? e->atv(e->P) : nullptr,
                                       training ? e->atv(e->GB) : nullptr, training ? e->at<uint64_t>(e->gate) : nullptr, st));
    else
      REGAT_TRY(regat_geoattn_fwd(dt, B, N, cf.nongt_dim, D, H, dirs, cf.pos_emb_dim, e->atv(e->Qb), e->atv(e->KVb), c.boxes, nullptr,
                                  e->wave_div, e->params + P0.v_off, wstride, ag, e->params + P0.b_off, bstride,
                                  e->at<float>(e->scal), e->atv(e->s), v0, cf.residual, e->atv(e->v1),
                                  training ? e->at<float>(e->P) : nullptr, training ? e->at<float>(e->GB) : nullptr,
                                  training ? e->at<uint64_t>(e->gate) : nullptr, st));
  }
  REGAT_TRY(fork_to(sd, st, e->ev[1]));   // join: weff / cb / uqe are ready
  REGAT_TRY(regat_butd_pool_fwd(dt, B, N, D, e->atv(e->v1), e->atv(e->weff), e->at<float>(e->cb), e->at<float>(e->att),
                                e->atv(e->pooled), st));
  if (att_out) REGAT_CUDA(cudaMemcpyAsync(att_out, e->atv(e->att), (size_t)B * N * sizeof(float), cudaMemcpyDeviceToDevice, st));
  // Training with the bf16 tensor-core kernels: the batch-sized products from here to the logits, the loss and their transposes
  // are the stages of one chained launch issued by backward() (csrc/gemm_tc.cu: gemm_chain_kernel)
  // (REGAT_CHAIN=1 opts in.  Measured at batch 256: the chain runs in 75 us against ~60 us for the separately launched kernels
  // inside the step's CUDA graph -- its stages keep only 24-98 CTAs busy and each streams 0.4-1.2 MB through one SM's L2 port,
  // while separate launches overlap their tails and the side stream's weight gradients; the step is 18 us slower with it.)
  const char* chain_s = getenv("REGAT_CHAIN");
  const int chain_env = chain_s ? atoi(chain_s) : 0;
  c.chain = training && e->grads && dt == REGAT_BF16 && e->use_tc && chain_env != 0 &&
            chain_fits(ceil_div(B, 128) * ceil_div(A, 64));
  if (c.chain) { c.logits_out = logits_out; return REGAT_OK; }
  REGAT_TRY(fc_fwd(e, st, e->l_ve, 0, B, D, e->atv(e->pooled), D, e->atv(e->pv), Hd, dt, false));
  REGAT_TRY(k_mul(dt, e->atv(e->pv), Hd, e->at<unsigned char>(e->uqe) + (size_t)Hd * es, 2 * Hd, e->atv(e->joint), Hd, B, Hd, st));
  // classifier                                                          classifier.py:14-25
  REGAT_TRY(fc_fwd(e, st, e->l_c0, 0, B, Hd, e->atv(e->joint), Hd, e->atv(e->hid), 2 * Hd, dt, true));
  // logits live in a buffer whose row pitch is padded to 16 bytes (3129 -> 3136) so the epilogue can use vector stores
  if (dt == REGAT_BF16 && e->use_tc && e->layers[e->l_c3].lowp_ld >= e->a_pad) {
    // all a_pad columns: the pad columns of the bf16 kernel copy and the floats behind the bias are zero (never written), so the
    // extra logits are 0 and the product has whole 32-column blocks -- the lean fp32 epilogue instead of the ragged generic one
    EpiArgs ep = epi0();
    ep.bias = biasp(e, e->l_c3);
    REGAT_TRY(dense(e, st, false, false, B, e->a_pad, 2 * Hd, e->atv(e->hid), 2 * Hd, W(e, e->l_c3), ldW(e, e->l_c3), e->atv(e->logits),
                    e->a_pad, REGAT_F32, ep));
  } else {
    REGAT_TRY(fc_fwd(e, st, e->l_c3, 0, B, 2 * Hd, e->atv(e->hid), 2 * Hd, e->atv(e->logits), e->a_pad, REGAT_F32, false));
  }
  if (logits_out)
    REGAT_CUDA(cudaMemcpy2DAsync(logits_out, (size_t)A * sizeof(float), e->atv(e->logits), (size_t)e->a_pad * sizeof(float),
                                 (size_t)A * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
  return REGAT_OK;
}

// Range r of the flat gradient buffer has received its last write: on the main stream and -- with side_work -- on the side
// stream (bias sums, the geometry reduction).
//   not fused: the side stream is joined into the main stream and the callback (library all-reduce driven from Python) is told;
//   fused train step: nothing is joined into the main stream -- the consumers wait for both producers by event.  Data parallel:
//   the `comm` stream exchanges the range over NVLink; then the `opt` stream runs clip + Adamax and re-derives alpha / the
//   bf16 kernels of the range (unless defer_opt: the caller queues optimize_range later).  Exchange r+1 overlaps optimizer r.
int range_ready(Ctx& c, int r, bool side_work, bool defer_opt = false) {
  regat_engine* e = c.e;
  const OptRange& R = e->rng[r];
  if (!c.fused_opt) {
    if (side_work) REGAT_TRY(fork_to(e->side, c.st, e->ev[16 + r]));
    if (e->grad_cb && R.l_last >= R.l_first) e->grad_cb(e->grad_cb_user, R.lo, R.hi - R.lo);
    return REGAT_OK;
  }
  const bool dp = e->dp_world > 1 && R.l_last >= R.l_first;
  cudaStream_t next = dp ? e->comm : e->opt;
  REGAT_TRY(fork_to(c.st, next, e->ev[10 + r]));
  if (side_work) REGAT_TRY(fork_to(e->side, next, e->ev[16 + r]));
  if (dp) {
    REGAT_TRY(dp_allreduce_f32_impl(e->dp_grad_ptrs, e->dp_mc, e->dp_flag_ptrs, e->dp_rank, e->dp_world, R.lo, R.hi - R.lo, 0,
                                    e->at<uint32_t>(e->hyp) + 8, e->dp_blocks, e->comm));
    REGAT_TRY(fork_to(e->comm, e->opt, e->ev[20 + r]));
  }
  if (!defer_opt) REGAT_TRY(optimize_range(e, r, e->opt));
  return REGAT_OK;
}

int backward(Ctx& c, const float* target, float grad_scale, float* dq_att, float* dq_last) {
  regat_engine* e = c.e;
  cudaStream_t st = c.st;
  const regat_config& cf = e->cfg;
  const int dt = e->dtype, B = c.B, N = c.N, M = c.M, R = c.R, Rm = c.Rm;
  const int V = cf.v_dim, Q = cf.q_dim, D = cf.rel_dim, H = cf.num_heads, A = cf.num_answers, Hd = cf.q_dim, dirs = cf.dir_num;
  const size_t es = dtype_size(dt);
  const void* feat = dt == REGAT_BF16 ? e->atv(e->featT) : (const void*)c.features;
  const void* qatt = dt == REGAT_BF16 ? e->atv(e->qattT) : (const void*)c.q_att;
  const void* qlast = dt == REGAT_BF16 ? e->atv(e->qlastT) : (const void*)c.q_last;
  const void* v0 = e->l_v2out >= 0 ? e->atv(e->v0) : feat;
  float* scal = e->at<float>(e->scal);
  e->grads_final = 0;

  cudaStream_t sd = e->side;
  unsigned char* dqe = e->at<unsigned char>(e->duqe) + (size_t)Hd * es;
  ColsumBatch cb_side, cb_qkv, cb_ds, cb_v0;   // bias gradients, one multi-problem launch per backward stage (all on the side stream)
  if (c.chain) {
    // pv -> joint -> hid -> logits -> loss / dlogits -> dhid -> djoint (dpv, dqe) -> dpooled as seven stages of ONE launch
    // (fusion.py:37-52, classifier.py:14-25, train.py:107-108 and their transposes); the weight gradients follow on the side stream
    ChainBuilder ch(e->at<unsigned int>(e->hyp) + 12);
    const unsigned char* qe = e->at<unsigned char>(e->uqe) + (size_t)Hd * es;
    ChainEpi ep;
    ep.bias = biasp(e, e->l_ve); ep.out2 = e->atv(e->joint); ep.out2_ld = Hd; ep.mul2 = qe; ep.mul2_ld = 2 * Hd;
    REGAT_TRY(ch.product(0, 0, B, Hd, D, e->atv(e->pooled), D, W(e, e->l_ve), ldW(e, e->l_ve), e->atv(e->pv), Hd, dt, ep));
    ep = ChainEpi(); ep.bias = biasp(e, e->l_c0); ep.relu = 1;
    REGAT_TRY(ch.product(0, 0, B, 2 * Hd, Hd, e->atv(e->joint), Hd, W(e, e->l_c0), ldW(e, e->l_c0), e->atv(e->hid), 2 * Hd, dt, ep));
    ep = ChainEpi(); ep.bias = biasp(e, e->l_c3);
    REGAT_TRY(ch.product(0, 0, B, A, 2 * Hd, e->atv(e->hid), 2 * Hd, W(e, e->l_c3), ldW(e, e->l_c3), e->atv(e->logits), e->a_pad, REGAT_F32, ep));
    REGAT_TRY(ch.loss(B, A, e->at<float>(e->logits), e->a_pad, target, grad_scale, scal + 2, scal + 3, e->atv(e->dlogits), e->a_pad));
    ep = ChainEpi(); ep.gate = e->atv(e->hid); ep.gate_ld = 2 * Hd;
    REGAT_TRY(ch.product(0, 1, B, 2 * Hd, A, e->atv(e->dlogits), e->a_pad, W(e, e->l_c3), ldW(e, e->l_c3), e->atv(e->dhid), 2 * Hd, dt, ep));
    ep = ChainEpi(); ep.mul = qe; ep.mul_ld = 2 * Hd; ep.out2 = dqe; ep.out2_ld = 2 * Hd; ep.mul2 = e->atv(e->pv); ep.mul2_ld = Hd;
    REGAT_TRY(ch.product(0, 1, B, Hd, 2 * Hd, e->atv(e->dhid), 2 * Hd, W(e, e->l_c0), ldW(e, e->l_c0), e->atv(e->dpv), Hd, dt, ep));
    ep = ChainEpi();
    REGAT_TRY(ch.product(0, 1, B, D, Hd, e->atv(e->dpv), Hd, W(e, e->l_ve), ldW(e, e->l_ve), e->atv(e->dpooled), D, dt, ep));
    {
      // one timing record for the whole chain: (M, N, K) = (B, 1, FLOPs / 2B) so that 2 M N K is the chain's FLOP count
      ProfScope prof(e, st, B, 1, 2 * (D * Hd + Hd * 2 * Hd + 2 * Hd * A));
      REGAT_TRY(ch.launch(st));
    }
    if (c.logits_out)
      REGAT_CUDA(cudaMemcpy2DAsync(c.logits_out, (size_t)A * sizeof(float), e->atv(e->logits), (size_t)e->a_pad * sizeof(float),
                                   (size_t)A * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
    // the classifier / BUTD weight gradients are issued further down, behind the attention backward pass (chain_wgrads)
  } else {
  // loss + dlogits                                                     train.py:107-108
  REGAT_TRY(k_bce(B, A, e->at<float>(e->logits), e->a_pad, target, grad_scale, scal + 2, scal + 3, e->atv(e->dlogits), e->a_pad, dt, st));
  // The input-gradient chain (dhid -> djoint -> dpv -> dpooled -> dv1) stays on the main stream; every weight / bias gradient of
  // the classifier and of BUTD is off the critical path and goes to the side stream as soon as its operands exist.
  REGAT_TRY(fork_to(st, sd, e->ev[2]));                    // dlogits ready
  if (dt == REGAT_BF16 && e->use_tc && (A % 4) != 0) {
    // the [2Hd, A] gradient has unaligned rows (A = 3129): compute it with a padded pitch, then compact into the flat buffer
    // (all a_pad columns: dlogits' pad columns are zero)
    REGAT_TRY(dense(e, sd, true, false, 2 * Hd, e->a_pad, B, e->atv(e->hid), 2 * Hd, e->atv(e->dlogits), e->a_pad, e->atv(e->dwc3), e->a_pad,
                    REGAT_F32, epi0()));
    REGAT_CUDA(cudaMemcpy2DAsync(gradW(e, e->l_c3), (size_t)A * sizeof(float), e->atv(e->dwc3), (size_t)e->a_pad * sizeof(float),
                                 (size_t)A * sizeof(float), 2 * Hd, cudaMemcpyDeviceToDevice, sd));
    REGAT_TRY(bias_grad(e, sd, e->atv(e->dlogits), e->a_pad, B, e->a_pad, gradB(e, e->l_c3), &cb_side));   // pad columns: zeros into the slot's alignment padding
  } else {
    REGAT_TRY(fc_wgrad(e, sd, e->l_c3, 0, B, 2 * Hd, e->atv(e->hid), 2 * Hd, e->atv(e->dlogits), e->a_pad, true, &cb_side));
  }
  REGAT_TRY(fc_dgrad(e, st, e->l_c3, 0, B, 2 * Hd, e->atv(e->dlogits), e->a_pad, e->atv(e->dhid), 2 * Hd, dt, false, e->atv(e->hid), 2 * Hd));
  REGAT_TRY(fork_to(st, sd, e->ev[3]));                    // dhid ready
  REGAT_TRY(fc_wgrad(e, sd, e->l_c0, 0, B, Hd, e->atv(e->joint), Hd, e->atv(e->dhid), 2 * Hd, true, &cb_side));
  REGAT_TRY(fc_dgrad(e, st, e->l_c0, 0, B, Hd, e->atv(e->dhid), 2 * Hd, e->atv(e->djoint), Hd, dt, false));
  // joint = pv * qe
  REGAT_TRY(k_mul_bwd(dt, e->atv(e->djoint), Hd, e->atv(e->pv), Hd, e->at<unsigned char>(e->uqe) + (size_t)Hd * es, 2 * Hd,
                      e->atv(e->dpv), Hd, dqe, 2 * Hd, B, Hd, st));
  REGAT_TRY(fork_to(st, sd, e->ev[4]));                    // dpv, dqe ready
  REGAT_TRY(fc_wgrad(e, sd, e->l_ve, 0, B, D, e->atv(e->pooled), D, e->atv(e->dpv), Hd, true, &cb_side));
  REGAT_TRY(fc_dgrad(e, st, e->l_ve, 0, B, D, e->atv(e->dpv), Hd, e->atv(e->dpooled), D, dt, false));
  }
  // attention pooling
  REGAT_TRY(regat_butd_pool_bwd(dt, B, N, D, e->atv(e->v1), e->atv(e->weff), e->at<float>(e->att), e->atv(e->dpooled),
                                e->atv(e->dv1), e->atv(e->dweff), e->at<float>(e->dcb), st));
  REGAT_TRY(fork_to(st, sd, e->ev[5]));                    // dweff, dcb ready: the rest of the question branch is side work
  {  // weff = alpha_va (uw Wva^T):  dWva_eff = dweff^T uw ;  duw = alpha_va (dweff Wva)
    EpiArgs ep = epi0();
    REGAT_TRY(dense(e, sd, true, false, D, Hd, B, e->atv(e->dweff), D, e->atv(e->uw), Hd, gradW(e, e->l_va), Hd, REGAT_F32, ep));
    ep.alpha = alpha_epi(e, e->l_va);
    REGAT_TRY(dense(e, sd, false, false, B, Hd, D, e->atv(e->dweff), D, W(e, e->l_va), ldW(e, e->l_va), e->atv(e->duw), Hd, dt, ep));
  }
  REGAT_TRY(k_butd_prep_bwd(dt, e->atv(e->duw), e->at<float>(e->dcb), e->atv(e->uqe), 2 * Hd, e->atv(e->uw),
                            e->params + e->layers[e->l_lin].v_off, alphap(e, e->l_lin), biasp(e, e->l_va), e->atv(e->duqe), 2 * Hd,
                            gradW(e, e->l_lin), gradB(e, e->l_va), gradB(e, e->l_lin), B, Hd, sd));
  if (dt == REGAT_BF16 && e->use_tc) {
    // [dW_qa | dW_qe] = q_last^T [du | dqe]  (one GEMM scattered into the two kernels' gradient slots);  dq_last = [du|dqe] [W_qa|W_qe]^T
    const long long offs[2] = {0, e->layers[e->l_qe].v_off - e->layers[e->l_qa].v_off};
    REGAT_TRY(dense_scatter(e, sd, Q, 2 * Hd, B, qlast, Q, e->atv(e->duqe), 2 * Hd, gradW(e, e->l_qa), Hd, Hd, offs));
    REGAT_TRY(bias_grad(e, sd, e->atv(e->duqe), 2 * Hd, B, Hd, gradB(e, e->l_qa), &cb_side));
    REGAT_TRY(bias_grad(e, sd, dqe, 2 * Hd, B, Hd, gradB(e, e->l_qe), &cb_side));
    if (dq_last)
      REGAT_TRY(dense(e, sd, false, true, B, Q, 2 * Hd, e->atv(e->duqe), 2 * Hd, lowp_at(e, e->guqe_off), 2 * Hd, dq_last, Q, REGAT_F32, epi0()));
  } else {
    REGAT_TRY(fc_wgrad(e, sd, e->l_qa, 0, B, Q, qlast, Q, e->atv(e->duqe), 2 * Hd, true, &cb_side));
    REGAT_TRY(fc_wgrad(e, sd, e->l_qe, 0, B, Q, qlast, Q, dqe, 2 * Hd, true, &cb_side));
    if (dq_last) {
      REGAT_TRY(fc_dgrad(e, sd, e->l_qa, 0, B, Q, e->atv(e->duqe), 2 * Hd, dq_last, Q, REGAT_F32, false));
      REGAT_TRY(fc_dgrad(e, sd, e->l_qe, 0, B, Q, dqe, 2 * Hd, dq_last, Q, REGAT_F32, true));
    }
  }
  if (!c.chain) {
    // BUTD + classifier gradients (the tail of the flat buffer) are final once the side stream has drained what is queued on it
    // now; nothing on the main stream writes them any more.  Announced BEFORE the attention backward pass is launched: that
    // kernel is latency-bound and uses little HBM bandwidth, the best partner the HBM-bound optimizer (and, data parallel, the
    // exchange) of this largest range can get.
    REGAT_TRY(k_colsum_multi(dt, cb_side, sd));
    REGAT_TRY(range_ready(c, 0, /*side_work=*/true));
  }
  // attention backward: dQ, dK, dV', dout (-> ds), dL (in place of P); then the geometry reduction
  if (attn_fast(e, N))
    REGAT_TRY(regat_attn_bwd_fast(B, N, cf.nongt_dim, D, H, dirs, e->atv(e->Qb), e->atv(e->KVb), e->atv(e->dv1), e->at<uint64_t>(e->gate),
                                  e->atv(e->P), e->atv(e->GB), e->atv(e->dQb), e->atv(e->dKVb), e->atv(e->ds), scal + 1, st));
  else
    REGAT_TRY(regat_attn_bwd(dt, B, N, cf.nongt_dim, D, H, dirs, e->atv(e->Qb), e->atv(e->KVb), e->atv(e->dv1),
                             e->at<uint64_t>(e->gate), e->at<float>(e->P), e->atv(e->dQb), e->atv(e->dKVb), e->atv(e->ds), st));
  if (c.chain) {
    // Weight gradients of the chained products: side stream, but only once the attention backward pass has run -- issued
    // earlier their CTAs take SM slots from that latency-bound kernel (measured: +20 us on the step)
    REGAT_TRY(fork_to(st, sd, e->ev[6]));                  // dQb, dKVb, dL final (and with them everything the chain produced)
    if ((A % 4) != 0) {
      REGAT_TRY(dense(e, sd, true, false, 2 * Hd, A, B, e->atv(e->hid), 2 * Hd, e->atv(e->dlogits), e->a_pad, e->atv(e->dwc3), e->a_pad,
                      REGAT_F32, epi0()));
      REGAT_CUDA(cudaMemcpy2DAsync(gradW(e, e->l_c3), (size_t)A * sizeof(float), e->atv(e->dwc3), (size_t)e->a_pad * sizeof(float),
                                   (size_t)A * sizeof(float), 2 * Hd, cudaMemcpyDeviceToDevice, sd));
      REGAT_TRY(bias_grad(e, sd, e->atv(e->dlogits), e->a_pad, B, e->a_pad, gradB(e, e->l_c3), &cb_side));   // pad columns: zeros into the slot's alignment padding
    } else {
      REGAT_TRY(fc_wgrad(e, sd, e->l_c3, 0, B, 2 * Hd, e->atv(e->hid), 2 * Hd, e->atv(e->dlogits), e->a_pad, true, &cb_side));
    }
    REGAT_TRY(fc_wgrad(e, sd, e->l_c0, 0, B, Hd, e->atv(e->joint), Hd, e->atv(e->dhid), 2 * Hd, true, &cb_side));
    REGAT_TRY(fc_wgrad(e, sd, e->l_ve, 0, B, D, e->atv(e->pooled), D, e->atv(e->dpv), Hd, true, &cb_side));
  }
  if (c.chain) {
    REGAT_TRY(k_colsum_multi(dt, cb_side, sd));
    REGAT_TRY(range_ready(c, 0, /*side_work=*/true));
  }
  // The bias gradients (column sums -- HBM-bound) and the small question-side products run on the side stream next to the
  // tensor-bound GEMMs of the main stream.  Gradient ranges are announced in the order they become final: attention layers,
  // then self_weights + label FC, then v2out -- the data-parallel layer starts each all-reduce behind the rest of the backward.
  if (!c.chain) REGAT_TRY(fork_to(st, sd, e->ev[6]));      // dQb, dKVb, dL final
  // the geometry reduction (SFU / issue bound, produces only dW_g, db_g, dc) runs beside the tensor-bound GEMMs
  {
    const Layer& P0 = e->layers[e->l_pos[0]];
    const long long wstride = dirs > 1 ? e->layers[e->l_pos[1]].v_off - P0.v_off : 0;
    const long long bstride = dirs > 1 ? e->layers[e->l_pos[1]].b_off - P0.b_off : 0;
    if (attn_fast(e, N))
      REGAT_TRY(regat_geo_bwd_fast(B, N, cf.nongt_dim, H, dirs, cf.pos_emb_dim, c.boxes, e->wave_div, e->atv(e->P), e->grads + P0.v_off,
                                   wstride, e->grads + P0.b_off, bstride, sd));
    else
      REGAT_TRY(regat_geo_bwd_ex(B, N, cf.nongt_dim, H, dirs, cf.pos_emb_dim, c.boxes, nullptr, e->wave_div, e->at<float>(e->P),
                                 e->at<float>(e->GB), e->grads + P0.v_off, wstride, e->grads + P0.b_off, bstride, scal + 1,
                                 dt == REGAT_BF16, sd));
    const Layer& LL = e->layers[e->l_label];
    REGAT_TRY(k_label_grad(scal + 1, e->grads, LL.v_off, LL.b_off, sd));
  }
  if (dt == REGAT_BF16 && e->use_tc) {
    for (int d = 0; d < dirs; ++d) {
      REGAT_TRY(bias_grad(e, sd, e->at<unsigned char>(e->dQb) + (size_t)d * D * es, dirs * D, R, D, gradB(e, e->l_q[d]), &cb_qkv));
      REGAT_TRY(bias_grad(e, sd, e->at<unsigned char>(e->dKVb) + (size_t)d * D * es, 2 * dirs * D, Rm, D, gradB(e, e->l_k[d]), &cb_qkv));
      REGAT_TRY(bias_grad(e, sd, e->at<unsigned char>(e->dKVb) + (size_t)(dirs + d) * D * es, 2 * dirs * D, Rm, D, gradB(e, e->l_out[d]), &cb_qkv));
    }
    REGAT_TRY(k_colsum_multi(dt, cb_qkv, sd));
    // weight gradients of side-by-side layers in one GEMM each (column blocks scattered to their own gradient slots),
    // input gradients with the direction / kind axis folded into K
    long long qo[2], kvo[4];
    for (int d = 0; d < dirs; ++d) {
      qo[d] = e->layers[e->l_q[d]].v_off - e->layers[e->l_q[0]].v_off;
      kvo[d] = e->layers[e->l_k[d]].v_off - e->layers[e->l_k[0]].v_off;
      kvo[dirs + d] = e->layers[e->l_out[d]].v_off - e->layers[e->l_k[0]].v_off;
    }
    REGAT_TRY(dense_scatter(e, st, D, dirs * D, R, e->atv(e->s), D, e->atv(e->dQb), dirs * D, gradW(e, e->l_q[0]), D, D, qo));
    REGAT_TRY(dense_scatter(e, st, D, 2 * dirs * D, Rm, e->atv(e->strunc), D, e->atv(e->dKVb), 2 * dirs * D, gradW(e, e->l_k[0]), D, D, kvo));
    // both attention layers: exchanged from here on; their optimizer step rewrites the bf16 kernels the two input-gradient
    // products below still read, so it is queued behind them
    REGAT_TRY(range_ready(c, 1, /*side_work=*/true, /*defer_opt=*/true));
    {
      // ds += dQ [W_q0|W_q1]^T and dstrunc = dKV [W_k0|W_k1|Kc_0|Kc_1]^T in ONE launch: 160 long tiles (K = 4 D) and 288 short ones
      // (K = 2 D) leave no part-filled wave between them
      ProfScope prof(e, st, R + Rm, D, (int)(((long long)R * dirs * D + (long long)Rm * 2 * dirs * D) / (R + Rm)));
      GemmCall ckv{0, 1, Rm, D, 2 * dirs * D, e->atv(e->dKVb), 2 * dirs * D, lowp_at(e, e->gkv_off), 2 * dirs * D, e->atv(e->dstrunc), D, dt, epi0()};
      GemmCall cq{0, 1, R, D, dirs * D, e->atv(e->dQb), dirs * D, lowp_at(e, e->gq_off), dirs * D, e->atv(e->ds), D, dt, epi0()};
      cq.e.accumulate = 1;                             // ds already holds dout
      REGAT_TRY(gemm_tc_pair(ckv, cq, st));
    }
    if (c.fused_opt) {
      REGAT_TRY(fork_to(st, e->opt, e->ev[14]));
      REGAT_TRY(optimize_range(e, 1, e->opt));
    }
  } else {
    for (int d = 0; d < dirs; ++d) {
      unsigned char* dQd = e->at<unsigned char>(e->dQb) + (size_t)d * D * es;
      unsigned char* dKd = e->at<unsigned char>(e->dKVb) + (size_t)d * D * es;
      unsigned char* dVd = e->at<unsigned char>(e->dKVb) + (size_t)(dirs + d) * D * es;
      REGAT_TRY(fc_wgrad(e, st, e->l_q[d], 0, R, D, e->atv(e->s), D, dQd, dirs * D, true, &cb_qkv));
      REGAT_TRY(fc_dgrad(e, st, e->l_q[d], 0, R, D, dQd, dirs * D, e->atv(e->ds), D, dt, true));
      REGAT_TRY(fc_wgrad(e, st, e->l_k[d], 0, Rm, D, e->atv(e->strunc), D, dKd, 2 * dirs * D, true, &cb_qkv));
      REGAT_TRY(fc_dgrad(e, st, e->l_k[d], 0, Rm, D, dKd, 2 * dirs * D, e->atv(e->dstrunc), D, dt, d > 0));
      REGAT_TRY(fc_wgrad(e, st, e->l_out[d], 0, Rm, D, e->atv(e->strunc), D, dVd, 2 * dirs * D, true, &cb_qkv));
      REGAT_TRY(fc_dgrad(e, st, e->l_out[d], 0, Rm, D, dVd, 2 * dirs * D, e->atv(e->dstrunc), D, dt, true));
    }
    REGAT_TRY(k_colsum_multi(dt, cb_qkv, sd));
    REGAT_TRY(range_ready(c, 1, /*side_work=*/true));  // every product that reads these layers' kernels has been issued
  }
  REGAT_TRY(k_addrows(dt, e->atv(e->ds), e->atv(e->dstrunc), B, N, M, D, st));
  // self_weights: s = alpha (v0 Ws[:D] + mask (q Ws[D:])) + b.   ds is final: its column sum and the masked segment sum go to
  // the side stream.  The order below puts the dependency chain first (dv0 -> v2out's weight gradient) and ends with the
  // smallest layer (self_weights, 1.8 M elements): that last range is the only exchange a data-parallel step cannot hide
  // behind compute.
  REGAT_TRY(fork_to(st, sd, e->ev[8]));
  REGAT_TRY(bias_grad(e, sd, e->atv(e->ds), D, R, D, gradB(e, e->l_self), &cb_ds));
  REGAT_TRY(k_colsum_multi(dt, cb_ds, sd));
  REGAT_TRY(k_segsum(dt, e->atv(e->ds), e->at<float>(e->mask), B, N, D, e->atv(e->dsq), sd));
  const Layer& LS = e->layers[e->l_self];
  if (e->l_v2out >= 0) {
    // dv0 = (dv1 [residual] + alpha ds Ws[:D]^T) o (v0 > 0), in place in the dv1 buffer
    REGAT_TRY(fc_dgrad(e, st, e->l_self, 0, R, D, e->atv(e->ds), D, e->atv(e->dv1), D, dt, cf.residual != 0, v0, D));
    REGAT_TRY(fork_to(st, sd, e->ev[2]));
    REGAT_TRY(bias_grad(e, sd, e->atv(e->dv1), D, R, D, gradB(e, e->l_v2out), &cb_v0));
    REGAT_TRY(k_colsum_multi(dt, cb_v0, sd));
    REGAT_TRY(fc_wgrad(e, st, e->l_v2out, 0, R, V, feat, V, e->atv(e->dv1), D, false));
  }
  // every announcement is a full join of the side stream (a data-parallel caller may end a CUDA-graph capture segment there)
  REGAT_TRY(fork_to(sd, st, e->ev[9]));
  REGAT_TRY(range_ready(c, 2, /*side_work=*/false));
  // self_weights last, as ONE range: an exchange has a fixed cost of ~40 us (three launches, two cross-GPU flag rounds), so
  // the tail of the step is one such exchange, not two
  REGAT_TRY(fc_wgrad(e, st, e->l_self, 0, R, D, v0, D, e->atv(e->ds), D, false));
  REGAT_TRY(fc_wgrad(e, st, e->l_self, D, B, Q, qatt, Q, e->atv(e->dsq), D, false));
  if (dq_att) REGAT_TRY(fc_dgrad(e, st, e->l_self, D, B, Q, e->atv(e->dsq), D, dq_att, Q, REGAT_F32, false));
  (void)LS;
  REGAT_TRY(range_ready(c, 3, /*side_work=*/false));   // self_weights (kernel, g, bias) and the label FC
  if (c.fused_opt) REGAT_TRY(fork_to(e->opt, st, e->ev[15]));      // the step is complete on the caller's stream
  return REGAT_OK;
}

int opt_stats(regat_engine* e, cudaStream_t st) {
  REGAT_TRY(prepare_weights(e, st));
  return k_opt_reduce(e->params, e->grads, e->tl_opt, e->chunks_opt, e->at<float>(e->partials), e->at<float>(e->stats), st,
                      e->at<unsigned int>(e->counters));
}

int check_call(regat_engine* e, int B, int N, bool need_opt) {
  REGAT_REQUIRE(e, REGAT_ERR_ARG, "engine is null");
  REGAT_REQUIRE(e->params && e->ws, REGAT_ERR_ARG, "engine: call regat_engine_bind first");
  REGAT_REQUIRE(!need_opt || (e->grads && e->am && e->au), REGAT_ERR_ARG, "engine: grads / Adamax buffers not bound");
  REGAT_REQUIRE(B > 0 && N > 0, REGAT_ERR_SHAPE, "engine: empty batch (B=%d, N=%d)", B, N);
  REGAT_REQUIRE(B <= e->max_b && N <= e->max_n, REGAT_ERR_SHAPE, "engine: batch %dx%d exceeds the capacity %dx%d given at create", B, N,
                e->max_b, e->max_n);
  return REGAT_OK;
}

}  // namespace
}  // namespace regat

extern "C" int regat_default_config(regat_config* cfg) {
  REGAT_REQUIRE(cfg, REGAT_ERR_ARG, "default_config: null pointer");
  cfg->v_dim = 2048; cfg->q_dim = 768; cfg->rel_dim = 1024; cfg->num_heads = 16; cfg->pos_emb_dim = 64; cfg->nongt_dim = 20;
  cfg->dir_num = 2; cfg->num_answers = 3129; cfg->label_bias = 0; cfg->residual = 1;
  cfg->grad_clip = 0.25f; cfg->beta1 = 0.9f; cfg->beta2 = 0.999f; cfg->eps = 1e-8f;
  return REGAT_OK;
}

extern "C" int regat_engine_create(const regat_config* cfg, int dtype, int max_batch, int max_rois, regat_engine** out) {
  REGAT_REQUIRE(cfg && out, REGAT_ERR_ARG, "engine_create: null pointer");
  REGAT_REQUIRE(dtype == REGAT_F32 || dtype == REGAT_BF16, REGAT_ERR_DTYPE, "engine_create: bad dtype %d", dtype);
  REGAT_REQUIRE(max_batch > 0 && max_rois > 0 && max_rois <= 128, REGAT_ERR_SHAPE, "engine_create: need 0 < max_rois <= 128, max_batch > 0");
  REGAT_REQUIRE(cfg->rel_dim == cfg->num_heads * 64, REGAT_ERR_UNSUPPORTED, "engine_create: head dim must be 64");
  REGAT_REQUIRE(cfg->pos_emb_dim == 64, REGAT_ERR_UNSUPPORTED, "engine_create: pos_emb_dim must be 64");
  REGAT_REQUIRE(cfg->dir_num >= 1 && cfg->dir_num <= 2, REGAT_ERR_SHAPE, "engine_create: dir_num must be 1 or 2 (graph_att_net.py:18)");
  REGAT_REQUIRE(cfg->v_dim % 8 == 0 && cfg->q_dim % 8 == 0 && cfg->nongt_dim > 0 && cfg->num_answers > 0, REGAT_ERR_SHAPE,
                "engine_create: v_dim and q_dim must be multiples of 8");
  regat_engine* e = new regat_engine();
  e->cfg = *cfg; e->dtype = dtype; e->max_b = max_batch; e->max_n = max_rois;
  const char* g = getenv("REGAT_GEMM");
  e->use_tc = !(g && strcmp(g, "simt") == 0);
  build_layout(e);
  e->ws_need = carve(e);
  build_lists(e);
  fill_wave_div(e->wave_div, cfg->pos_emb_dim);
  *out = e;
  return REGAT_OK;
}
extern "C" int regat_engine_destroy(regat_engine* e) {
  if (e) {
    for (int i = 0; i < regat_engine::NEV; ++i) if (e->ev[i]) cudaEventDestroy(e->ev[i]);
    if (e->side) cudaStreamDestroy(e->side);
    if (e->opt) cudaStreamDestroy(e->opt);
    if (e->comm) cudaStreamDestroy(e->comm);
    delete e;
  }
  return REGAT_OK;
}

extern "C" int regat_engine_sizes(const regat_engine* e, int64_t* param_elems, int64_t* workspace_bytes) {
  REGAT_REQUIRE(e, REGAT_ERR_ARG, "engine is null");
  if (param_elems) *param_elems = e->param_elems;
  if (workspace_bytes) *workspace_bytes = e->ws_need;
  return REGAT_OK;
}
extern "C" int regat_engine_param(const regat_engine* e, int idx, int64_t* offset, int64_t* numel, int32_t* layer, int32_t* kind) {
  REGAT_REQUIRE(e, REGAT_ERR_ARG, "engine is null");
  if (idx < 0 || idx >= (int)e->entries.size()) return REGAT_ERR_ARG;
  const Entry& x = e->entries[idx];
  if (offset) *offset = x.off;
  if (numel) *numel = x.numel;
  if (layer) *layer = x.layer;
  if (kind) *kind = x.kind;
  return REGAT_OK;
}
extern "C" int regat_engine_bind(regat_engine* e, float* params, float* grads, float* adamax_m, float* adamax_u, void* workspace,
                                 int64_t workspace_bytes) {
  REGAT_REQUIRE(e && params && workspace, REGAT_ERR_ARG, "engine_bind: null pointer");
  REGAT_REQUIRE(workspace_bytes >= e->ws_need, REGAT_ERR_WORKSPACE, "engine_bind: workspace %lld B < required %lld B",
                (long long)workspace_bytes, e->ws_need);
  REGAT_REQUIRE(((uintptr_t)workspace & 255) == 0 && ((uintptr_t)params & 255) == 0 && (!grads || ((uintptr_t)grads & 255) == 0),
                REGAT_ERR_ALIGN, "engine_bind: buffers must be 256-byte aligned");
  e->params = params; e->grads = grads; e->am = adamax_m; e->au = adamax_u;
  e->ws = static_cast<unsigned char*>(workspace); e->ws_bytes = workspace_bytes;
  e->weights_ready = 0;
  REGAT_CUDA(cudaMemset(e->ws + e->counters.off, 0, MAX_TENSORS * 4));
  REGAT_CUDA(cudaMemset(e->ws + e->hyp.off, 0, 64 * 4));    // setup-time: the counters reset themselves afterwards
  return REGAT_OK;
}

extern "C" int regat_engine_forward(regat_engine* e, int B, int N, const float* features, const float* boxes, const float* q_att,
                                    const float* q_last, float* logits, float* att, regat_stream_t stream) {
  REGAT_TRY(check_call(e, B, N, false));
  REGAT_REQUIRE(features && boxes && q_att && q_last, REGAT_ERR_ARG, "engine_forward: null input");
  const int l0 = launch_counter();
  Ctx c{e, (cudaStream_t)stream, B, N, std::min(e->cfg.nongt_dim, N), B * N, B * std::min(e->cfg.nongt_dim, N), features, boxes, q_att, q_last};
  REGAT_TRY(forward(c, false, logits, att));
  e->last_launches = launch_counter() - l0;
  return REGAT_OK;
}

extern "C" int regat_engine_fwd_bwd(regat_engine* e, int B, int N, const float* features, const float* boxes, const float* q_att,
                                    const float* q_last, const float* target, float grad_scale, float* loss_out, float* logits,
                                    float* dq_att, float* dq_last, regat_stream_t stream) {
  REGAT_TRY(check_call(e, B, N, false));
  REGAT_REQUIRE(e->grads, REGAT_ERR_ARG, "engine_fwd_bwd: grads buffer not bound");
  REGAT_REQUIRE(features && boxes && q_att && q_last && target, REGAT_ERR_ARG, "engine_fwd_bwd: null input");
  const int l0 = launch_counter();
  cudaStream_t st = (cudaStream_t)stream;
  Ctx c{e, st, B, N, std::min(e->cfg.nongt_dim, N), B * N, B * std::min(e->cfg.nongt_dim, N), features, boxes, q_att, q_last};
  REGAT_TRY(forward(c, true, logits, nullptr));
  REGAT_TRY(backward(c, target, grad_scale, dq_att, dq_last));
  if (loss_out) REGAT_CUDA(cudaMemcpyAsync(loss_out, e->at<float>(e->scal) + 2, 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  e->last_launches = launch_counter() - l0;
  return REGAT_OK;
}

// Converts the grads buffer in place from dL/dW_eff to the reference's (dv, dg) -- what tape.gradient returns
// (train.py:111).  Optional: regat_engine_update accepts either state.
extern "C" int regat_engine_finalize_grads(regat_engine* e, regat_stream_t stream) {
  REGAT_REQUIRE(e && e->grads && e->ws, REGAT_ERR_ARG, "engine_finalize_grads: not bound");
  if (e->grads_final) return REGAT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  REGAT_TRY(opt_stats(e, st));
  REGAT_TRY(k_opt_finalize(e->params, e->grads, e->tl_opt, e->chunks_opt, e->at<float>(e->stats), e->at<float>(e->alpha),
                           e->at<float>(e->invn), st));
  e->grads_final = 1;
  return REGAT_OK;
}

extern "C" int regat_engine_update(regat_engine* e, float lr, int step, regat_stream_t stream) {
  REGAT_REQUIRE(e && e->params && e->grads && e->am && e->au && e->ws, REGAT_ERR_ARG, "engine_update: buffers not bound");
  REGAT_REQUIRE(step >= 1, REGAT_ERR_ARG, "engine_update: step is 1-based");
  cudaStream_t st = (cudaStream_t)stream;
  const int l0 = launch_counter();
  REGAT_TRY(prepare_weights(e, st));          // alpha / ||v|| of the CURRENT parameters (no-op unless they were written from outside)
  REGAT_TRY(k_hyper(e->at<Hyper>(e->hyp), lr, step - 1, e->cfg.beta1, HYPER_SET_LR | HYPER_SET_STEP | HYPER_TICK, st));
  REGAT_TRY(optimize_range(e, -1, st));       // clip + Adamax, then alpha / bf16 copies of the new parameters
  e->last_launches = launch_counter() - l0;
  return REGAT_OK;
}

extern "C" int regat_engine_set_lr(regat_engine* e, float lr, regat_stream_t stream) {
  REGAT_REQUIRE(e && e->ws, REGAT_ERR_ARG, "engine_set_lr: not bound");
  return k_hyper(e->at<Hyper>(e->hyp), lr, 0, e->cfg.beta1, HYPER_SET_LR, (cudaStream_t)stream);
}
extern "C" int regat_engine_set_step(regat_engine* e, int steps_done, regat_stream_t stream) {
  REGAT_REQUIRE(e && e->ws && steps_done >= 0, REGAT_ERR_ARG, "engine_set_step: not bound / negative step count");
  return k_hyper(e->at<Hyper>(e->hyp), 0.f, steps_done, e->cfg.beta1, HYPER_SET_STEP, (cudaStream_t)stream);
}
extern "C" int regat_engine_get_step(regat_engine* e, int* steps_done, float* lr, regat_stream_t stream) {
  REGAT_REQUIRE(e && e->ws, REGAT_ERR_ARG, "engine_get_step: not bound");
  Hyper h;
  REGAT_CUDA(cudaMemcpyAsync(&h, e->at<Hyper>(e->hyp), sizeof(h), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  REGAT_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  if (steps_done) *steps_done = h.step;
  if (lr) *lr = h.lr;
  return REGAT_OK;
}

extern "C" int regat_engine_train_step(regat_engine* e, int B, int N, const float* features, const float* boxes, const float* q_att,
                                       const float* q_last, const float* target, float lr, int step, float* loss_out,
                                       regat_stream_t stream) {
  REGAT_TRY(check_call(e, B, N, true));
  REGAT_REQUIRE(step >= 1, REGAT_ERR_ARG, "engine_train_step: step is 1-based");
  const int l0 = launch_counter();
  REGAT_TRY(k_hyper(e->at<Hyper>(e->hyp), lr, step - 1, e->cfg.beta1, HYPER_SET_LR | HYPER_SET_STEP, (cudaStream_t)stream));
  REGAT_TRY(regat_engine_train_step_dev(e, B, N, features, boxes, q_att, q_last, target, loss_out, stream));
  e->last_launches = launch_counter() - l0;
  return REGAT_OK;
}

// One whole train step from device-resident optimizer state: no host scalar enters any launch, so the call can be captured
// in ONE CUDA graph and replayed (also with the in-place gradient exchange of regat_engine_set_dp).
extern "C" int regat_engine_train_step_dev(regat_engine* e, int B, int N, const float* features, const float* boxes, const float* q_att,
                                           const float* q_last, const float* target, float* loss_out, regat_stream_t stream) {
  REGAT_TRY(check_call(e, B, N, true));
  REGAT_REQUIRE(features && boxes && q_att && q_last && target, REGAT_ERR_ARG, "engine_train_step: null input");
  const int l0 = launch_counter();
  cudaStream_t st = (cudaStream_t)stream;
  Ctx c{e, st, B, N, std::min(e->cfg.nongt_dim, N), B * N, B * std::min(e->cfg.nongt_dim, N), features, boxes, q_att, q_last};
  c.fused_opt = true;
  REGAT_TRY(forward(c, true, nullptr, nullptr));
  REGAT_TRY(backward(c, target, 1.0f / (float)e->dp_world, nullptr, nullptr));
  if (loss_out) REGAT_CUDA(cudaMemcpyAsync(loss_out, e->at<float>(e->scal) + 2, 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  e->last_launches = launch_counter() - l0;
  return REGAT_OK;
}

// Data parallel: the bound grads buffer is a symmetric allocation (same layout on every rank, peer-mapped, optionally bound to an
// NVSwitch multicast address); regat_engine_train_step[_dev] then reduces every gradient range in place over NVLink on the
// engine's own stream as soon as the backward pass has finished it, with grad_scale = 1/world.  world = 1 switches it off.
extern "C" int regat_engine_set_dp(regat_engine* e, const uint64_t* grad_ptrs, uint64_t multicast_ptr, const uint64_t* flag_ptrs,
                                   int rank, int world, int blocks) {
  REGAT_REQUIRE(e, REGAT_ERR_ARG, "engine is null");
  if (world <= 1) { e->dp_world = 1; e->dp_rank = 0; pdl_enabled() = 1; return REGAT_OK; }
  // data parallel: no programmatic dependent launch (process-wide, one engine per rank): with it the dependents of the tcgen05 and
  // optimizer kernels become resident early and delay the exchange kernels -- measured on 8 GPUs 1.025 -> 0.986 ms per step
  // without it (REGAT_DP_PDL=1 keeps it on)
  { const char* s = getenv("REGAT_DP_PDL"); pdl_enabled() = (s && atoi(s) != 0) ? 1 : 0; }
  REGAT_REQUIRE(grad_ptrs && flag_ptrs && world <= 16 && rank >= 0 && rank < world, REGAT_ERR_ARG, "engine_set_dp: bad rank %d / world %d", rank, world);
  REGAT_REQUIRE(e->grads && (uint64_t)(uintptr_t)e->grads == grad_ptrs[rank], REGAT_ERR_ARG,
                "engine_set_dp: the bound grads buffer must be this rank's entry of the symmetric pointer table");
  for (int r = 0; r < world; ++r) { e->dp_grad_ptrs[r] = grad_ptrs[r]; e->dp_flag_ptrs[r] = flag_ptrs[r]; }
  e->dp_mc = multicast_ptr; e->dp_rank = rank; e->dp_world = world; e->dp_blocks = blocks > 0 ? blocks : 32;
  REGAT_CUDA(cudaMemset(e->at<uint32_t>(e->hyp) + 8, 0, sizeof(uint32_t)));      // exchange call counter: every rank starts at 0
  return REGAT_OK;
}

// Derives alpha, the bf16 kernels, the gathered biases and the label constant from the parameters NOW (on `stream`).  Call it
// after writing the parameter buffer from outside when captured graphs of engine calls are going to be replayed next.
extern "C" int regat_engine_refresh_weights(regat_engine* e, regat_stream_t stream) {
  REGAT_REQUIRE(e && e->params && e->ws, REGAT_ERR_ARG, "engine_refresh_weights: not bound");
  e->weights_ready = 0;
  return prepare_weights(e, (cudaStream_t)stream);
}

// Measurement aid: with on != 0 every dense product of the following EAGER engine calls is bracketed by timing events (never
// enable it while capturing a graph); regat_engine_profile_read synchronises the device and returns (M, N, K, ms) per product
// in launch order, then clears the list.  Products run beside the side-stream work they overlap with in the real step.
extern "C" int regat_engine_profile(regat_engine* e, int on) {
  REGAT_REQUIRE(e, REGAT_ERR_ARG, "engine is null");
  for (auto& r : e->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  e->prof.clear();
  e->prof_on = on != 0;
  return REGAT_OK;
}
extern "C" int regat_engine_profile_read(regat_engine* e, int max_records, int32_t* mnk, float* ms, int* count) {
  REGAT_REQUIRE(e && mnk && ms && count, REGAT_ERR_ARG, "engine_profile_read: null pointer");
  REGAT_CUDA(cudaDeviceSynchronize());
  int n = 0;
  for (auto& r : e->prof) {
    if (n < max_records) {
      float t = 0.f;
      REGAT_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
      mnk[3 * n] = r.M; mnk[3 * n + 1] = r.N; mnk[3 * n + 2] = r.K; ms[n] = t;
      ++n;
    }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  e->prof.clear();
  *count = n;
  return REGAT_OK;
}

extern "C" int regat_engine_last_launches(const regat_engine* e) { return e ? e->last_launches : 0; }

// The caller has written the parameter buffer (checkpoint load, set_weights, ...): cached weight-norm statistics are stale.
extern "C" int regat_engine_params_changed(regat_engine* e) {
  REGAT_REQUIRE(e, REGAT_ERR_ARG, "engine is null");
  e->weights_ready = 0;
  return REGAT_OK;
}

extern "C" int regat_engine_set_grad_callback(regat_engine* e, regat_grad_ready_fn fn, void* user) {
  REGAT_REQUIRE(e, REGAT_ERR_ARG, "engine is null");
  e->grad_cb = fn; e->grad_cb_user = user;
  return REGAT_OK;
}

extern "C" int regat_engine_set_wave_div(regat_engine* e, const float* wave_div_host) {
  REGAT_REQUIRE(e && wave_div_host, REGAT_ERR_ARG, "engine_set_wave_div: null pointer");
  for (int k = 0; k < 8; ++k) e->wave_div[k] = wave_div_host[k];
  return REGAT_OK;
}

extern "C" int regat_engine_config(const regat_engine* e, regat_config* cfg) {
  REGAT_REQUIRE(e && cfg, REGAT_ERR_ARG, "engine_config: null pointer");
  *cfg = e->cfg;
  return REGAT_OK;
}

// raw views of internal activations for tests / the Python layer mirror (name lookup keeps the ABI small)
extern "C" int regat_engine_buffer(const regat_engine* e, const char* name, void** ptr) {
  REGAT_REQUIRE(e && name && ptr && e->ws, REGAT_ERR_ARG, "engine_buffer: null / unbound");
#define REGAT_BUF(n) if (strcmp(name, #n) == 0) { *ptr = e->ws + e->n.off; return REGAT_OK; }
  REGAT_BUF(v0) REGAT_BUF(mask) REGAT_BUF(s) REGAT_BUF(Qb) REGAT_BUF(KVb) REGAT_BUF(v1) REGAT_BUF(P) REGAT_BUF(GB) REGAT_BUF(att)
  REGAT_BUF(pooled) REGAT_BUF(joint) REGAT_BUF(hid) REGAT_BUF(logits) REGAT_BUF(alpha) REGAT_BUF(scal) REGAT_BUF(dv1) REGAT_BUF(ds)
  REGAT_BUF(dQb) REGAT_BUF(dKVb)
#undef REGAT_BUF
  REGAT_REQUIRE(false, REGAT_ERR_ARG, "engine_buffer: unknown buffer '%s'", name);
}
