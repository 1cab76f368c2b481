// Shared device/host helpers for the lb-wavenet B200 kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <mma.h>

#include <cstdint>
#include <cstdio>

#include "model.h"

namespace wn {

typedef __nv_bfloat16 bf16;

extern int64_t g_launches;  // kernels launched by this library (bench.py gpu_launches)

#define WN_CUDA_CHECK(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      wn::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return WN_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

#define WN_LAUNCH_CHECK()                                                                       \
  do {                                                                                          \
    ++wn::g_launches;                                                                           \
    cudaError_t _e = cudaGetLastError();                                                        \
    if (_e != cudaSuccess) {                                                                    \
      wn::set_error("%s:%d: kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return WN_ERR_CUDA;                                                                       \
    }                                                                                           \
  } while (0)

// ---- per-category kernel timing (CUDA events on the launch stream; bench.py's roofline) ------
enum ProfCat {
  PROF_PREP = 0, PROF_LAYER_FWD = 1, PROF_POST_FWD = 2, PROF_POST_BWD = 3, PROF_LAYER_BWD_A = 4,  // 4: fused layer backward (or gate backward on the GC path)
  PROF_LAYER_BWD_B = 5, PROF_WGRAD = 6, PROF_EMBED_GC_BWD = 7, PROF_ADAM = 8, PROF_GEN = 9, PROF_NCAT = 16
};
struct ProfScope {
  int cat;
  cudaStream_t st;
  bool on;
  size_t rec = 0;  // index of this scope's record (scopes may nest: the LC backward runs inside the PRE / GC scope)
  ProfScope(int cat, cudaStream_t st);
  ~ProfScope();
};

// One event pair around a run of launches of the same category (the per-layer loops): events between the launches
// would serialise them and defeat programmatic dependent launch.  ProfScope(cat) inside the group is a no-op.
struct ProfGroup {
  int cat;
  cudaStream_t st;
  bool on;
  ProfGroup(int cat, cudaStream_t st, bool enable = true);
  ~ProfGroup();
};

// ---- programmatic dependent launch (sm_90+): the next kernel of the stream may be scheduled while this one drains ----
// pdl_launch_dependents(): "my dependents may be scheduled now" (they still block in pdl_wait()).
// pdl_wait(): every prerequisite grid has completed and its memory is visible.  No-ops without the launch attribute.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Launch with programmatic stream serialisation: the kernel may be scheduled before its predecessor in the stream has
// drained (it blocks in pdl_wait()), which hides the ~4 us launch gap between the 60 per-layer launches of a step.
inline bool pdl_enabled() {
  static const bool on = getenv("WN_DISABLE_PDL") == nullptr;
  return on;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, int block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

// ---- optional in-kernel timeline (tools/trace_layer.py): CTA 0's role warps log (event, tile, clock64) ----
extern int g_trace_layer;       // layer whose launches are traced (-1: all)
extern long long* g_trace_buf;  // device buffer [32 warps][WN_TRACE_PER_WARP] or nullptr (set by wn_debug_trace)
constexpr int WN_TRACE_PER_WARP = 2048;
struct Tracer {
  long long* p;
  int n;
  __device__ __forceinline__ void init(long long* base, int warp, bool on) {
    p = (base != nullptr && on) ? base + (size_t)warp * WN_TRACE_PER_WARP : nullptr;
    n = 0;
  }
  __device__ __forceinline__ void ev(int code, int tile) {
    if (p != nullptr && n < WN_TRACE_PER_WARP)
      p[n++] = ((long long)code << 56) | ((long long)(tile & 0xffff) << 40) | (clock64() & 0xffffffffffLL);
  }
};

// ---- global loads that stay where they are written -------------------------------------------------
// 16-byte read-only global load (asm volatile: the compiler may not move it), zero when !pred
__device__ __forceinline__ uint4 ldg_nc_v4_pred(const void* p, bool pred) {
  uint4 v;
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.u32 q, %5, 0;\n"
      "mov.u32 %0, 0;\n"
      "mov.u32 %1, 0;\n"
      "mov.u32 %2, 0;\n"
      "mov.u32 %3, 0;\n"
      "@q ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];\n"
      "}\n"
      : "=&r"(v.x), "=&r"(v.y), "=&r"(v.z), "=&r"(v.w)
      : "l"(p), "r"((uint32_t)pred));
  return v;
}
__device__ __forceinline__ uint4 ldg_nc_v4_pinned(const void* p) { return ldg_nc_v4_pred(p, true); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// The same interface compiled to nothing: the post-net and weight-gradient kernels log only when the library is built
// with -DWN_POST_TRACE (tools/trace_layer.py postfwd | postbwd | wgrad); left in, the hooks cost their single-thread
// producer / issuer loops ~0.02 ms per step each.
struct NoTracer {
  __device__ __forceinline__ void init(long long*, int, bool) {}
  __device__ __forceinline__ void ev(int, int) {}
};
#ifdef WN_POST_TRACE
using PostTracer = Tracer;
#else
using PostTracer = NoTracer;
#endif
// the per-layer kernels (layer_umma.cu) likewise: -DWN_LAYER_TRACE for tools/trace_layer.py fwd | bwd (measured with
// the hooks compiled out: the 30 backward launches 1.33 -> 1.29 ms per configs[1] step, the forward unchanged)
#ifdef WN_LAYER_TRACE
using LayerTracer = Tracer;
constexpr bool kLayerTrace = true;
#else
using LayerTracer = NoTracer;
constexpr bool kLayerTrace = false;
#endif

// ---- math ------------------------------------------------------------------------------
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return 0.5f * tanh_fast(0.5f * x) + 0.5f; }

__device__ __forceinline__ float bf2f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ bf16 f2bf(float v) { return __float2bfloat16_rn(v); }

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

// ---- legacy-tensor-path tile GEMM (nvcuda::wmma, HMMA) ---------------------------------
// C_s[TM x N] (=|+=) A_s[TM x K] * B[K x N];  A_s bf16 in shared memory (row-major, lda),
// B bf16 in GLOBAL memory: row-major [K][N] (ldb = row length) or, with COLB, stored as
// B^T row-major [N][K] (ldb = K-extent of the stored matrix).  All warps of the CTA take
// part; caller synchronises before and after.
template <int TM, bool COLB, bool ACCUM>
__device__ __forceinline__ void tile_mma(const bf16* A_s, int lda, const bf16* __restrict__ Bg,
                                         int ldb, float* C_s, int ldc, int K, int N) {
  using namespace nvcuda;
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int ntn = N >> 4;
  const int ntiles = (TM / 16) * ntn;
  for (int tile = warp; tile < ntiles; tile += nwarps) {
    const int mi = tile % (TM / 16), ni = tile / (TM / 16);
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc;
    float* cptr = C_s + mi * 16 * ldc + ni * 16;
    if (ACCUM)
      wmma::load_matrix_sync(acc, cptr, ldc, wmma::mem_row_major);
    else
      wmma::fill_fragment(acc, 0.f);
    for (int k0 = 0; k0 < K; k0 += 16) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, bf16, wmma::row_major> fa;
      wmma::load_matrix_sync(fa, A_s + mi * 16 * lda + k0, lda);
      if (COLB) {
        wmma::fragment<wmma::matrix_b, 16, 16, 16, bf16, wmma::col_major> fb;
        wmma::load_matrix_sync(fb, Bg + (size_t)(ni * 16) * ldb + k0, ldb);
        wmma::mma_sync(acc, fa, fb, acc);
      } else {
        wmma::fragment<wmma::matrix_b, 16, 16, 16, bf16, wmma::row_major> fb;
        wmma::load_matrix_sync(fb, Bg + (size_t)k0 * ldb + ni * 16, ldb);
        wmma::mma_sync(acc, fa, fb, acc);
      }
    }
    wmma::store_matrix_sync(cptr, acc, ldc, wmma::mem_row_major);
  }
}

// copy `ncols` bf16 (multiple of 8) of one global row into shared memory, 16 B at a time,
// executed by the calling thread for chunk index c
__device__ __forceinline__ void copy16(bf16* dst, const bf16* src, bool valid) {
  uint4 v = make_uint4(0, 0, 0, 0);
  if (valid) v = *reinterpret_cast<const uint4*>(src);
  *reinterpret_cast<uint4*>(dst) = v;
}

}  // namespace wn
