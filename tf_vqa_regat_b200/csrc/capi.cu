// extern "C" glue: error strings, device queries, the generic dense entry, weight-norm entries and the
// DLPack front door.  See include/regat.h for the contract.
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"
#include "kernels.h"

namespace regat {

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
  return REGAT_ERR_CUDA;
}
int& launch_counter() { return g_launches; }
int& pdl_enabled() { static int on = 1; return on; }

int num_sms() {
  static int cached = 0;
  if (!cached) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

}  // namespace regat

using namespace regat;

extern "C" int regat_abi_version(void) { return REGAT_ABI_VERSION; }

extern "C" int regat_last_error(char* buf, size_t n) {
  const size_t len = strlen(g_err);
  if (buf && n) {
    const size_t k = len < n - 1 ? len : n - 1;
    memcpy(buf, g_err, k);
    buf[k] = 0;
  }
  return (int)len;
}

extern "C" int regat_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// Host <-> device copies and a device-wide synchronisation for bindings whose tensor provider exposes neither (TensorFlow eager):
// thin wrappers over the CUDA runtime the library links statically, so such a binding needs no other CUDA dependency.
extern "C" int regat_memcpy(void* dst, const void* src, int64_t bytes, int kind, regat_stream_t stream) {
  REGAT_REQUIRE(dst && src && bytes >= 0, REGAT_ERR_ARG, "memcpy: null pointer / negative size");
  REGAT_REQUIRE(kind >= 1 && kind <= 3, REGAT_ERR_ARG, "memcpy: kind must be 1 (host to device), 2 (device to host) or 3 (device to device)");
  if (bytes == 0) return REGAT_OK;
  const cudaMemcpyKind k = kind == 1 ? cudaMemcpyHostToDevice : (kind == 2 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
  REGAT_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, k, (cudaStream_t)stream));
  if (kind != 3) REGAT_CUDA(cudaStreamSynchronize((cudaStream_t)stream));      // host memory is safe to reuse / read on return
  return REGAT_OK;
}
extern "C" int regat_device_synchronize(void) {
  REGAT_CUDA(cudaDeviceSynchronize());
  return REGAT_OK;
}

extern "C" int regat_gemm_trace(void* device_buf) {
  regat::gemm_tc_set_trace(static_cast<long long*>(device_buf));
  return REGAT_OK;
}

extern "C" int regat_gemm(int dtype, int transA, int transB, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
                          void* C, int ldc, int c_dtype, const regat_epilogue* epi, regat_stream_t stream) {
  REGAT_REQUIRE(A && B && C, REGAT_ERR_ARG, "gemm: null pointer");
  REGAT_REQUIRE(dtype == REGAT_F32 || dtype == REGAT_BF16, REGAT_ERR_DTYPE, "gemm: bad dtype %d", dtype);
  REGAT_REQUIRE(c_dtype == REGAT_F32 || c_dtype == dtype, REGAT_ERR_DTYPE, "gemm: C must be fp32 or the activation dtype");
  REGAT_REQUIRE(M >= 0 && N >= 0 && K > 0, REGAT_ERR_SHAPE, "gemm: bad shape %dx%dx%d", M, N, K);
  REGAT_REQUIRE(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, REGAT_ERR_SHAPE, "gemm: leading dimension too small");
  if (regat_device_count() == 0) { set_error("gemm: no CUDA device (there is no CPU fallback)"); return REGAT_ERR_CUDA; }
  EpiArgs e;
  memset(&e, 0, sizeof(e));
  int split_k = 1;
  if (epi) {
    e.alpha = epi->alpha; e.alpha_cols = epi->alpha_cols; e.bias = epi->bias;
    e.addend = epi->addend; e.addend_ld = epi->addend_ld; e.addend_rows = epi->addend_rows; e.row_scale = epi->row_scale;
    e.relu = epi->relu; e.accumulate = epi->accumulate; e.gate = epi->gate; e.gate_ld = epi->gate_ld;
    e.c2 = epi->c2; e.c2_ld = epi->c2_ld; e.c2_rows_in = epi->c2_rows_in; e.c2_rows_keep = epi->c2_rows_keep;
    split_k = epi->split_k > 1 ? epi->split_k : 1;
    REGAT_REQUIRE(!e.addend || e.addend_rows > 0, REGAT_ERR_ARG, "gemm: addend_rows must be positive");
    REGAT_REQUIRE(!e.c2 || (e.c2_rows_in > 0 && e.c2_rows_keep > 0), REGAT_ERR_ARG, "gemm: bad c2 row mapping");
  }
  cudaStream_t st = (cudaStream_t)stream;
  const char* g = getenv("REGAT_GEMM");
  const bool force_simt = g && strcmp(g, "simt") == 0;
  if (dtype == REGAT_BF16 && !force_simt) {
    REGAT_REQUIRE(gemm_tc_supported(transA, transB, M, N, K, A, lda, B, ldb), REGAT_ERR_ALIGN,
                  "gemm(bf16): operands must be 16-byte aligned with leading dimensions that are multiples of 8");
    return gemm_tc(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, c_dtype, e, split_k, st);
  }
  return gemm_simt(dtype, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, c_dtype, e, st);
}

static int fill_list(TensorList& tl, const int64_t* off, const int64_t* numel, const int32_t* cols, const int64_t* g_off,
                     const int64_t* off_lowp, const int32_t* ld_lowp, int n) {
  REGAT_REQUIRE(n > 0 && n <= MAX_TENSORS, REGAT_ERR_SHAPE, "weight norm: between 1 and %d tensors per call", MAX_TENSORS);
  memset(&tl, 0, sizeof(tl));
  tl.n = n;
  for (int i = 0; i < n; ++i) {
    tl.off[i] = off ? off[i] : 0; tl.numel[i] = numel ? numel[i] : 0; tl.cols[i] = cols ? cols[i] : 1;
    tl.g_off[i] = g_off ? g_off[i] : 0; tl.off_lowp[i] = off_lowp ? off_lowp[i] : 0; tl.ld_lowp[i] = ld_lowp ? ld_lowp[i] : tl.cols[i];
    tl.layer[i] = i;
  }
  return REGAT_OK;
}

extern "C" int regat_wn_prepare(const float* params, const int64_t* v_off_host, const int64_t* v_numel_host,
                                const int32_t* v_cols_host, int n_layers, float* sumsq, void* v_lowp,
                                const int64_t* off_lowp_host, const int32_t* ld_lowp_host, regat_stream_t stream) {
  REGAT_REQUIRE(params && v_off_host && v_numel_host && v_cols_host && sumsq, REGAT_ERR_ARG, "wn_prepare: null pointer");
  REGAT_REQUIRE(!v_lowp || (off_lowp_host && ld_lowp_host), REGAT_ERR_ARG, "wn_prepare: bf16 copy needs offsets and leading dims");
  TensorList tl;
  REGAT_TRY(fill_list(tl, v_off_host, v_numel_host, v_cols_host, nullptr, off_lowp_host, ld_lowp_host, n_layers));
  const int chunks = build_tensor_list(tl);
  return k_wn_prepare(params, tl, chunks, sumsq, v_lowp, (cudaStream_t)stream);
}

extern "C" int regat_wn_alpha(const float* params, const int64_t* g_off_host, int n_layers, const float* sumsq, float* alpha,
                              float* inv_norm, regat_stream_t stream) {
  REGAT_REQUIRE(params && g_off_host && sumsq && alpha && inv_norm, REGAT_ERR_ARG, "wn_alpha: null pointer");
  REGAT_REQUIRE(n_layers <= 32, REGAT_ERR_SHAPE, "wn_alpha: at most 32 layers per call");
  TensorList tl;
  REGAT_TRY(fill_list(tl, nullptr, nullptr, nullptr, g_off_host, nullptr, nullptr, n_layers));
  return k_wn_alpha(params, tl, const_cast<float*>(sumsq), alpha, inv_norm, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ DLPack front door
// Minimal restatement of the DLPack v0.x ABI (dlpack.h, struct DLManagedTensor); only what is read here.
namespace {
typedef enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3 } DLDeviceType;
typedef struct { int32_t device_type; int32_t device_id; } DLDevice;
typedef struct { uint8_t code; uint8_t bits; uint16_t lanes; } DLDataType;   // code 2 = float
typedef struct {
  void* data; DLDevice device; int32_t ndim; DLDataType dtype; int64_t* shape; int64_t* strides; uint64_t byte_offset;
} DLTensor;
}  // namespace
struct DLManagedTensor { DLTensor dl_tensor; void* manager_ctx; void (*deleter)(struct DLManagedTensor*); };

static int dl_f32(struct DLManagedTensor* m, const char* name, int ndim, const int64_t* want, float** out) {
  REGAT_REQUIRE(m, REGAT_ERR_ARG, "%s: null DLManagedTensor", name);
  const DLTensor& t = m->dl_tensor;
  int dev = 0;
  cudaGetDevice(&dev);
  REGAT_REQUIRE(t.device.device_type == kDLCUDA && t.device.device_id == dev, REGAT_ERR_DEVICE,
                "%s: tensor must live on CUDA device %d (got type %d id %d)", name, dev, t.device.device_type, t.device.device_id);
  REGAT_REQUIRE(t.dtype.code == 2 && t.dtype.bits == 32 && t.dtype.lanes == 1, REGAT_ERR_DTYPE, "%s: tensor must be float32", name);
  REGAT_REQUIRE(t.ndim == ndim, REGAT_ERR_SHAPE, "%s: expected %d dims, got %d", name, ndim, t.ndim);
  int64_t stride = 1;
  for (int i = ndim - 1; i >= 0; --i) {
    REGAT_REQUIRE(want[i] < 0 || t.shape[i] == want[i], REGAT_ERR_SHAPE, "%s: dim %d is %lld, expected %lld", name, i,
                  (long long)t.shape[i], (long long)want[i]);
    REGAT_REQUIRE(!t.strides || t.shape[i] == 1 || t.strides[i] == stride, REGAT_ERR_SHAPE, "%s: tensor must be compact row-major", name);
    stride *= t.shape[i];
  }
  *out = reinterpret_cast<float*>(static_cast<char*>(t.data) + t.byte_offset);
  REGAT_REQUIRE(aligned16(*out), REGAT_ERR_ALIGN, "%s: data pointer not 16-byte aligned", name);
  return REGAT_OK;
}

extern "C" int regat_engine_config(const regat_engine* e, regat_config* cfg);

static int dl_inputs(regat_engine* e, struct DLManagedTensor* features, struct DLManagedTensor* boxes, struct DLManagedTensor* q_att,
                     struct DLManagedTensor* q_last, int* B, int* N, float** f, float** bx, float** qa, float** ql) {
  REGAT_REQUIRE(e && features, REGAT_ERR_ARG, "engine / features is null");
  regat_config cfg;
  REGAT_TRY(regat_engine_config(e, &cfg));
  REGAT_REQUIRE(features->dl_tensor.ndim == 3, REGAT_ERR_SHAPE, "features: expected [B,N,v_dim]");
  const int64_t b = features->dl_tensor.shape[0], n = features->dl_tensor.shape[1];
  const int64_t sf[3] = {b, n, cfg.v_dim}, sb[3] = {b, n, 4}, sq[2] = {b, cfg.q_dim};
  REGAT_TRY(dl_f32(features, "features", 3, sf, f));
  REGAT_TRY(dl_f32(boxes, "boxes", 3, sb, bx));
  REGAT_TRY(dl_f32(q_att, "q_att", 2, sq, qa));
  REGAT_TRY(dl_f32(q_last, "q_last", 2, sq, ql));
  *B = (int)b; *N = (int)n;
  return REGAT_OK;
}

extern "C" int regat_engine_forward_dl(regat_engine* e, struct DLManagedTensor* features, struct DLManagedTensor* boxes,
                                       struct DLManagedTensor* q_att, struct DLManagedTensor* q_last,
                                       struct DLManagedTensor* logits_out, regat_stream_t stream) {
  int B, N; float *f, *bx, *qa, *ql, *lo;
  REGAT_TRY(dl_inputs(e, features, boxes, q_att, q_last, &B, &N, &f, &bx, &qa, &ql));
  regat_config cfg;
  REGAT_TRY(regat_engine_config(e, &cfg));
  const int64_t sl[2] = {B, cfg.num_answers};
  REGAT_TRY(dl_f32(logits_out, "logits_out", 2, sl, &lo));
  return regat_engine_forward(e, B, N, f, bx, qa, ql, lo, nullptr, stream);
}

extern "C" int regat_engine_train_step_dl(regat_engine* e, struct DLManagedTensor* features, struct DLManagedTensor* boxes,
                                          struct DLManagedTensor* q_att, struct DLManagedTensor* q_last,
                                          struct DLManagedTensor* target, float lr, int step, float* loss_out,
                                          regat_stream_t stream) {
  int B, N; float *f, *bx, *qa, *ql, *tg;
  REGAT_TRY(dl_inputs(e, features, boxes, q_att, q_last, &B, &N, &f, &bx, &qa, &ql));
  regat_config cfg;
  REGAT_TRY(regat_engine_config(e, &cfg));
  const int64_t st[2] = {B, cfg.num_answers};
  REGAT_TRY(dl_f32(target, "target", 2, st, &tg));
  return regat_engine_train_step(e, B, N, f, bx, qa, ql, tg, lr, step, loss_out, stream);
}
