// Internal launch API shared by pointwise.cu, engine.cu and capi.cu.
#pragma once
#include "common.cuh"

namespace regat {

constexpr int MAX_TENSORS = 40;

// A list of tensors inside the flat fp32 parameter buffer, passed to kernels by value.
struct TensorList {
  int n;
  long long off[MAX_TENSORS];       // element offset of the tensor in params / grads / adamax slots
  long long numel[MAX_TENSORS];
  long long g_off[MAX_TENSORS];     // offset of the scalar g of the owning weight-normed layer
  long long off_lowp[MAX_TENSORS];  // element offset of the bf16 copy
  int ld_lowp[MAX_TENSORS];
  int cols[MAX_TENSORS];            // last dimension
  int kind[MAX_TENSORS];            // 0 = weight-normed kernel v, 1 = bias
  int layer[MAX_TENSORS];           // index into alpha / inv_norm
  int chunk_start[MAX_TENSORS + 1]; // first block handling each tensor (filled by build_tensor_list)
  int vchunk_start[MAX_TENSORS];    // kind-0 tensors of the optimizer list: first chunk of the same tensor in the weight-norm list
};
int build_tensor_list(TensorList& tl);  // returns number of blocks

struct OptHyper {
  float lr_t;       // lr / (1 - beta1^t)
  float beta1, beta2, eps, clip;
  int grads_are_final;
  const float* lr_t_dev = nullptr;   // when set, lr_t is read from device memory (written by k_hyper: CUDA-graph capturable steps)
};

// Device-resident optimizer scalars (engine workspace): lr, the number of optimizer steps taken, lr_t of the current step.
struct Hyper { float lr; int step; float lr_t; int pad; };
constexpr int HYPER_SET_LR = 1, HYPER_SET_STEP = 2, HYPER_TICK = 4;
// one thread: optionally overwrite lr / step, then (TICK) ++step and lr_t = lr / (1 - beta1^step)   (Keras Adamax, train.py:48)
int k_hyper(Hyper* h, float lr, int step, float beta1, int mode, cudaStream_t st);
__device__ __forceinline__ void hyper_tick(Hyper* h, float beta1) {
  const int t = ++h->step;
  h->lr_t = (float)((double)h->lr / (1.0 - pow((double)beta1, (double)t)));
}

int k_wn_prepare(const float* params, const TensorList& tl, int chunks, float* sumsq, void* lowp, cudaStream_t st, float* partials = nullptr);
int k_wn_scaled_copy(const float* params, const TensorList& tl, int chunks, const float* alpha, void* lowp, cudaStream_t st);
int k_gather(const float* src, const TensorList& tl, float* dst, cudaStream_t st);
int k_wn_alpha(const float* params, const TensorList& tl, float* sumsq, float* alpha, float* inv_norm, cudaStream_t st, const float* partials = nullptr,
               const TensorList* gather = nullptr, float* gather_out = nullptr, int label_layer = -1, long long label_v_off = 0,
               long long label_b_off = -1, float* label_c = nullptr);
int k_cast(int to_dtype, const float* in, void* out, long long n, cudaStream_t st);
int k_rowmask(int dt, const void* v, int rows, int D, float* mask, cudaStream_t st);
int k_mul(int dt, const void* a, int lda, const void* b, int ldb, void* out, int ldo, int rows, int cols, cudaStream_t st);
int k_mul_bwd(int dt, const void* dz, int ldz, const void* a, int lda, const void* b, int ldb, void* da, int ldda, void* db,
              int lddb, int rows, int cols, cudaStream_t st);
int k_butd_prep(int dt, const void* u, int ldu, const float* vl, const float* alpha_l, const float* bva, const float* bl,
                void* uw, float* cb, int B, int Hd, cudaStream_t st);
int k_butd_prep_bwd(int dt, const void* duw, const float* dcb, const void* u, int ldu, const void* uw, const float* vl,
                    const float* alpha_l, const float* bva, void* du, int lddu, float* dwl, float* dbva, float* dbl, int B,
                    int Hd, cudaStream_t st);
int k_bce(int B, int A, const float* logits, int ldl, const float* target, float gscale, float* loss, float* score,
          void* dlog, int ldd, int d_dtype, cudaStream_t st);
constexpr int COLSUM_SLAB = 128;     // rows per block of the batched column-sum kernel
struct ColsumBatch {
  int n = 0;
  const void* x[12]; float* out[12];
  int ld[12], rows[12], cols[12];
  int slab_start[13];
  void add(const void* xp, int ldp, int r, int c, float* o) { if (o && r > 0) { x[n] = xp; ld[n] = ldp; rows[n] = r; cols[n] = c; out[n] = o; ++n; } }
};
int k_colsum_multi(int dt, ColsumBatch& cb, cudaStream_t st);
int k_colsum(int dt, const void* x, int ld, int rows, int cols, float* out, cudaStream_t st);
int k_segsum(int dt, const void* x, const float* w, int B, int N, int D, void* out, cudaStream_t st);
int k_addrows(int dt, void* dst, const void* src, int B, int N, int M, int D, cudaStream_t st);
int k_opt_reduce(const float* params, const float* grads, const TensorList& tl, int chunks, float* partials, float* stats, cudaStream_t st,
                 unsigned int* counters = nullptr);
int k_opt_update(float* params, const float* grads, float* m, float* u, const TensorList& tl, int chunks, const float* stats,
                 const float* alpha, const float* inv_norm, const OptHyper& hp, cudaStream_t st, float* vpartials = nullptr);
int k_opt_finalize(const float* params, float* grads, const TensorList& tl, int chunks, const float* stats, const float* alpha,
                   const float* inv_norm, cudaStream_t st);
int k_label_const(const float* params, long long v_off, long long b_off, const float* alpha_l, float* c, cudaStream_t st);
int k_label_grad(const float* dc, float* grads, long long v_off, long long b_off, cudaStream_t st);

}  // namespace regat
