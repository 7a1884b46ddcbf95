// Shared helpers for libregat.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/regat.h"

namespace regat {

typedef __nv_bfloat16 bf16;

// ---- host-side error plumbing (thread-local message, integer status across the ABI) ----
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
int& launch_counter();  // thread-local count of kernel launches issued through the library

#define REGAT_CUDA(expr)                                                        \
  do {                                                                          \
    cudaError_t _e = (expr);                                                    \
    if (_e != cudaSuccess) return ::regat::cuda_fail(_e, #expr, __FILE__, __LINE__); \
  } while (0)

#define REGAT_REQUIRE(cond, code, ...)     \
  do {                                     \
    if (!(cond)) {                         \
      ::regat::set_error(__VA_ARGS__);     \
      return (code);                       \
    }                                      \
  } while (0)

// after every <<<>>> : count it and surface launch-configuration errors immediately
#define REGAT_POST_LAUNCH()          \
  do {                               \
    ++::regat::launch_counter();     \
    REGAT_CUDA(cudaGetLastError());  \
  } while (0)

#define REGAT_TRY(expr)              \
  do {                               \
    int _s = (expr);                 \
    if (_s != REGAT_OK) return _s;   \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t round_up64(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
inline size_t dtype_size(int dt) { return dt == REGAT_BF16 ? 2 : 4; }

int num_sms();  // cached multiprocessor count of the current device
// Programmatic dependent launch for the library's kernel chains (tcgen05 products, optimizer kernels).  On by default; the engine
// switches it off for data-parallel steps (measured on 8 GPUs: early-resident dependents delay the exchange kernels, +40 us).
int& pdl_enabled();

// ---- device-side scalar conversions ----
__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(bf16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- the one epilogue every GEMM kernel applies (see regat.h: regat_epilogue) ----
struct EpiArgs {
  const float* alpha; int alpha_cols;
  const float* bias;
  const float* addend; int addend_ld; int addend_rows; const float* row_scale;
  int relu;
  int accumulate;
  const void* gate; int gate_ld;
  void* c2; int c2_ld; int c2_rows_in; int c2_rows_keep;
};

// TC = storage type of C / C2 / gate.  Scalar form; callers vectorise around it.
template <typename TC>
__device__ __forceinline__ float epi_value(const EpiArgs& e, int r, int c, float x, const TC* C, int ldc) {
  if (e.addend) x += (e.row_scale ? e.row_scale[r] : 1.f) * e.addend[(size_t)(r / e.addend_rows) * e.addend_ld + c];
  if (e.alpha) x *= e.alpha[e.alpha_cols ? c / e.alpha_cols : 0];
  if (e.bias) x += e.bias[c];
  if (e.relu) x = fmaxf(x, 0.f);
  if (e.accumulate) x += to_f(C[(size_t)r * ldc + c]);
  if (e.gate && !(to_f(reinterpret_cast<const TC*>(e.gate)[(size_t)r * e.gate_ld + c]) > 0.f)) x = 0.f;
  return x;
}
template <typename TC>
__device__ __forceinline__ void epi_store(const EpiArgs& e, int r, int c, float x, TC* C, int ldc) {
  x = epi_value<TC>(e, r, c, x, C, ldc);
  C[(size_t)r * ldc + c] = from_f<TC>(x);
  if (e.c2) {
    int rr = r % e.c2_rows_in;
    if (rr < e.c2_rows_keep)
      reinterpret_cast<TC*>(e.c2)[((size_t)(r / e.c2_rows_in) * e.c2_rows_keep + rr) * e.c2_ld + c] = from_f<TC>(x);
  }
}

// ---- internal entry points shared between translation units ----
int gemm_simt(int in_dtype, int transA, int transB, int M, int N, int K, const void* A, int lda,
              const void* B, int ldb, void* C, int ldc, int c_dtype, const EpiArgs& e, cudaStream_t st);
int gemm_tc(int transA, int transB, int M, int N, int K, const void* A, int lda, const void* B, int ldb,
            void* C, int ldc, int c_dtype, const EpiArgs& e, int split_k, cudaStream_t st, int c_block_cols = 0,
            const long long* c_block_off = nullptr, int max_ctas = 0);
struct GemmCall {
  int transA, transB, M, N, K;
  const void* A; int lda; const void* B; int ldb; void* C; int ldc; int c_dtype;
  EpiArgs e;
};
int gemm_tc_pair(const GemmCall& c0, const GemmCall& c1, cudaStream_t st);
// Dependent chain of batch-sized products (+ the loss) as stages of ONE persistent launch (csrc/gemm_tc.cu, gemm_chain_kernel).
struct ChainEpi {
  const float* bias = nullptr; int relu = 0;
  const void* gate = nullptr; int gate_ld = 0;          // bf16: keep where gate > 0
  const void* mul = nullptr; int mul_ld = 0;            // bf16: C = x * mul
  void* out2 = nullptr; int out2_ld = 0; const void* mul2 = nullptr; int mul2_ld = 0;   // bf16: out2 = x * mul2
};
class ChainBuilder {
 public:
  explicit ChainBuilder(unsigned int* counter);          // device counter, zero between launches
  ~ChainBuilder();
  ChainBuilder(const ChainBuilder&) = delete;
  ChainBuilder& operator=(const ChainBuilder&) = delete;
  int product(int transA, int transB, int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* C, int ldc, int c_dtype,
              const ChainEpi& e);
  int loss(int B, int A, const float* logits, int ldl, const float* target, float gscale, float* loss, float* score, void* dlog, int ldd);
  int launch(cudaStream_t st);
 private:
  struct Impl;
  Impl* impl;
};
bool chain_fits(int max_units);
void gemm_tc_set_trace(long long* buf);
bool gemm_tc_supported(int transA, int transB, int M, int N, int K, const void* A, int lda, const void* B,
                       int ldb);

}  // namespace regat
