// Training hot path, generation 1: fused per-stage kernels with an HMMA (nvcuda::wmma) tile
// GEMM inside.  Same fusion structure, buffers and rounding points as the tcgen05 kernels in
// train_umma.cu, which replace the dominant contractions; these remain the reference device
// implementation for shapes the tcgen05 kernels do not cover.
//
// Reference semantics: tmodel.py:292-328 (forward), tmodel.py:218-261 (loss),
// tmodel.py:354-358 (gradients), train.py:178,186 (Adam).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

#include "common.cuh"

namespace wn {

int64_t g_launches = 0;
long long* g_trace_buf = nullptr;
int g_trace_layer = -1;

// ---- per-category event timing ------------------------------------------------------------
struct ProfState {
  uint32_t mask = 0;  // bit c: category c is timed
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  struct Rec { int cat; size_t e0, e1; int64_t launches; };
  std::vector<Rec> recs;
};
static ProfState g_prof;

static size_t prof_event(cudaStream_t st) {
  if (g_prof.used == g_prof.pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    g_prof.pool.push_back(e);
  }
  cudaEventRecord(g_prof.pool[g_prof.used], st);
  return g_prof.used++;
}

static int g_prof_group_cat = -1;  // category of the open ProfGroup, -1: none

ProfScope::ProfScope(int cat_, cudaStream_t st_)
    : cat(cat_), st(st_), on(((g_prof.mask >> cat_) & 1u) && g_prof_group_cat != cat_) {
  if (on) {
    rec = g_prof.recs.size();
    g_prof.recs.push_back({cat, prof_event(st), 0, g_launches});
  }
}
ProfGroup::ProfGroup(int cat_, cudaStream_t st_, bool enable)
    : cat(cat_), st(st_), on(enable && ((g_prof.mask >> cat_) & 1u) && g_prof_group_cat < 0) {
  if (on) {
    g_prof.recs.push_back({cat, prof_event(st), 0, g_launches});
    g_prof_group_cat = cat;
  }
}
ProfGroup::~ProfGroup() {
  if (on) {
    g_prof_group_cat = -1;
    for (size_t i = g_prof.recs.size(); i-- > 0;)
      if (g_prof.recs[i].cat == cat && g_prof.recs[i].e1 == 0) {  // the group's own record (inner scopes of other categories follow it)
        g_prof.recs[i].e1 = prof_event(st);
        g_prof.recs[i].launches = g_launches - g_prof.recs[i].launches;
        break;
      }
  }
}
ProfScope::~ProfScope() {
  if (on && rec < g_prof.recs.size()) {
    g_prof.recs[rec].e1 = prof_event(st);
    g_prof.recs[rec].launches = g_launches - g_prof.recs[rec].launches;
  }
}

// tcgen05 generation (train_umma.cu)
bool umma_post_supported(const wn_model* m);
bool umma_post_chain_supported(const wn_model* m);  // GEMM-chain post-net (n_skip / n_post up to 512)
int launch_post_fwd_chain_umma(wn_model* m, const float* d_params, unsigned char* ws, const int32_t* d_wav,
                               const int32_t* d_ids, int T, double* d_stats, float* d_logits, cudaStream_t st);
int launch_post_bwd_chain_umma(wn_model* m, unsigned char* ws, int T, float* d_grads, cudaStream_t st);
bool umma_wide_layer_supported(const wn_model* m);  // layers with R, D multiples of 64 through the GEMM kernel
int launch_prep_wide_umma(wn_model* m, const float* d_params, unsigned char* ws, cudaStream_t st);
int launch_layer_fwd_wide_umma(wn_model* m, const float* d_params, unsigned char* ws, int T, int l, cudaStream_t st);
int launch_layer_bwd_wide_umma(wn_model* m, const float* d_params, unsigned char* ws, int T, int l, float* d_grads,
                               cudaStream_t st);
bool umma_wgrad_x_supported(const wn_model* m, int T);
int launch_wgrad_umma_x(wn_model* m, const bf16* xfull, int dil, int T, int tap, const bf16* Y, int ldy, int N,
                        float* out, float* out2, int n_split, int ldo, int64_t tap_out_stride, cudaStream_t st);
int launch_wgrad_umma_cols(wn_model* m, const bf16* A, int lda, int M_total, const bf16* Y, int N_total, int64_t rows,
                           float* out, int mode, float* grads, cudaStream_t st);
int launch_prep_umma(wn_model* m, const float* d_params, unsigned char* ws, cudaStream_t st);
bool umma_layer_supported(const wn_model* m);
int launch_prep_layer_umma(wn_model* m, const float* d_params, unsigned char* ws, cudaStream_t st);
int launch_layer_fwd_umma(wn_model* m, const float* d_params, unsigned char* ws, const int32_t* d_ids, int T, int l,
                          cudaStream_t st);
bool umma_bwd_fused_supported(const wn_model* m);
int launch_layer_bwd_fused_umma(wn_model* m, const float* d_params, unsigned char* ws, const int32_t* d_ids, int T, int l,
                                float* d_grads, cudaStream_t st);
bool umma_wgrad_supported(const wn_model* m, int lda, int ldy, int N);
int launch_wgrad_umma(wn_model* m, const bf16* A, int lda, int a_col0, int M_total, const bf16* Y, int ldy, int N,
                      int64_t rows, float* out, int ldo, int mode, float* grads, cudaStream_t st);
int launch_post_bwd_umma(wn_model* m, unsigned char* ws, int T, float* d_grads, cudaStream_t st);
int launch_lc_fwd(wn_model* m, const float* d_params, const float* d_mel, unsigned char* ws, int T, cudaStream_t st);
int launch_lc_bwd(wn_model* m, unsigned char* ws, int T, float* d_grads, cudaStream_t st);
int launch_post_fwd_umma(wn_model* m, const float* d_params, unsigned char* ws, const int32_t* d_wav,
                         const int32_t* d_ids, int T, double* d_stats, float* d_logits, cudaStream_t st);

constexpr int TM = 64;   // timesteps per CTA tile
constexpr int NT = 256;  // threads per CTA

struct Dims {
  int B, T, R, D, S, P, Q, L, LD, G, C1, use_bias;
  int64_t rows;
};

// ======================================================================================
// parameter preparation
// ======================================================================================
__global__ void k_cast_params(const float* __restrict__ p, bf16* __restrict__ w, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) w[i] = f2bf(p[i]);
}

// out[s] = sum_l SKIP_BIAS_l[s]   (bias of the concatenated-K skip GEMM)
__global__ void k_skip_bias_sum(const float* __restrict__ p, const LayerDesc* __restrict__ layers,
                                int L, int S, float* __restrict__ out) {
  // the layers' offsets first (in parallel), then independent loads: as written before -- offset, then value, layer after
  // layer in one thread -- this was 2 L dependent global-memory round trips (~12 us for 30 layers)
  __shared__ int64_t off_s[256];
  for (int l = threadIdx.x; l < L && l < 256; l += blockDim.x) off_s[l] = layers[l].skip_b;
  __syncthreads();
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float acc = 0.f;
  for (int l = 0; l < L; ++l) {
    const int64_t o = l < 256 ? off_s[l] : layers[l].skip_b;
    if (o >= 0) acc += p[o + s];
  }
  out[s] = acc;
}

// tbl[l][c][n] = sum_g GC_EMBED[c][g] * (n<D ? GC_SIGNAL_l[g][n] : GC_GATE_l[g][n-D])
// (tmodel.py:112-113,150-154 folded into a per-(layer, voice id) additive table)
__global__ void k_gc_table(const float* __restrict__ p, int64_t off_embed,
                           const LayerDesc* __restrict__ layers, int C1, int G, int D,
                           float* __restrict__ tbl) {
  const int l = blockIdx.y, c = blockIdx.x, n = threadIdx.x;
  if (n >= 2 * D) return;
  const float* e = p + off_embed + (int64_t)c * G;
  const float* w = n < D ? p + layers[l].gc_sig + n : p + layers[l].gc_gate + (n - D);
  float acc = 0.f;
  for (int g = 0; g < G; ++g) acc += e[g] * w[(int64_t)g * D];
  tbl[((int64_t)l * C1 + c) * 2 * D + n] = acc;
}

// ======================================================================================
// PRE gather (one_hot @ PRE == row gather; tmodel.py:53-66,96-100) and SAVE prefix handling
// ======================================================================================
// one thread per 8 channels (one 16-byte store); R % 8 == 0 (the registry enforces R % 16 == 0)
__global__ void k_embed(const float* __restrict__ p, int64_t off_pre, int64_t off_pre_b,
                        const int32_t* __restrict__ wav, bf16* __restrict__ x0, int B, int T, int R,
                        int dil0, int Q) {
  const int R8 = R >> 3;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * T * R8) return;
  const int r = (int)(i % R8) * 8;
  const int64_t bt = i / R8;
  const int t = (int)(bt % T), b = (int)(bt / T);
  const int code = __ldg(wav + bt);
  float v[8];
  if (code >= 0 && code < Q) {  // tf.one_hot: an out-of-range index gives an all-zero row
    const float4 a = __ldg(reinterpret_cast<const float4*>(p + off_pre + (int64_t)code * R + r));
    const float4 c = __ldg(reinterpret_cast<const float4*>(p + off_pre + (int64_t)code * R + r) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c.x; v[5] = c.y; v[6] = c.z; v[7] = c.w;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = 0.f;
  }
  if (off_pre_b >= 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] += __ldg(p + off_pre_b + r + k);
  }
  __nv_bfloat162 o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) o[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
  *reinterpret_cast<uint4*>(x0 + ((int64_t)b * (dil0 + T) + dil0 + t) * R + r) = *reinterpret_cast<uint4*>(o);
}

// xfull_l[b][j][:] = SAVE_l[b][j][:]  for j < dil     (the concat of tmodel.py:127)
__global__ void k_save_load(const bf16* __restrict__ save, unsigned char* __restrict__ ws,
                            const LayerDesc* __restrict__ layers, int B, int T, int R) {
  const LayerDesc ld = layers[blockIdx.y];
  const int64_t n = (int64_t)B * ld.dil * R;
  const bf16* src = save + ld.save_off;
  bf16* dst = reinterpret_cast<bf16*>(ws + ld.xfull_off);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / ((int64_t)ld.dil * R), rem = i % ((int64_t)ld.dil * R);
    dst[b * (int64_t)(ld.dil + T) * R + rem] = src[i];
  }
}

// SAVE_l[b][j][:] = full_l[b][T + j][:]  (last dil rows of [SAVE; x], tmodel.py:165; covers T < dil)
__global__ void k_save_store(bf16* __restrict__ save, const unsigned char* __restrict__ ws,
                             const LayerDesc* __restrict__ layers, int B, int T, int R) {
  const LayerDesc ld = layers[blockIdx.y];
  const int64_t n = (int64_t)B * ld.dil * R;
  bf16* dst = save + ld.save_off;
  const bf16* src = reinterpret_cast<const bf16*>(ws + ld.xfull_off);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / ((int64_t)ld.dil * R), rem = i % ((int64_t)ld.dil * R);
    dst[i] = src[b * (int64_t)(ld.dil + T) * R + (int64_t)T * R + rem];
  }
}

// ======================================================================================
// layer forward: dilated conv (both taps, SIGNAL|GATE) + bias + GC + gate + RESIDUAL 1x1 + add
// (tmodel.py:117-168, 171-184, 325)
// ======================================================================================
struct LayerArgs {
  const bf16* wbf;
  const float* params;
  LayerDesc ld;
  const bf16* xin;   // xfull_l   [B][dil+T][R]
  bf16* xout;        // xfull_{l+1} or nullptr for the last layer
  int dil_next;
  bf16* z;           // [B*T][LD]
  const float* gc_tbl;  // this layer's [C1][2D] table or nullptr
  const int32_t* ids;
  int B, T, R, D, LD, l, C1;
};

__device__ __forceinline__ void load_conv_tile(bf16* A_s, int lda, const bf16* xb, int t0, int T,
                                               int dil, int R) {
  const int cpr = 2 * R / 8, half_c = R / 8;
  for (int idx = threadIdx.x; idx < TM * cpr; idx += NT) {
    const int r = idx / cpr, c = idx % cpr;
    const int half = c >= half_c;
    const int cc = c - half * half_c;
    const int t = t0 + r;
    const bf16* src = xb + (size_t)(t + (half ? dil : 0)) * R + cc * 8;
    copy16(A_s + r * lda + c * 8, src, t < T);
  }
}

__global__ void __launch_bounds__(NT) k_layer_fwd(LayerArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int R = a.R, D = a.D;
  const int lda = 2 * R + 8, ldc = 2 * D + 4, ldz = D + 8, ldc2 = R + 4;
  bf16* A_s = reinterpret_cast<bf16*>(smem);
  float* C_s = reinterpret_cast<float*>(A_s + TM * lda);
  bf16* Z_s = reinterpret_cast<bf16*>(C_s + TM * ldc);
  float* C2_s = reinterpret_cast<float*>(Z_s + TM * ldz);
  const int b = blockIdx.y, t0 = blockIdx.x * TM, T = a.T, dil = a.ld.dil;
  const bf16* xb = a.xin + (size_t)b * (dil + T) * R;
  load_conv_tile(A_s, lda, xb, t0, T, dil, R);
  __syncthreads();
  tile_mma<TM, false, false>(A_s, lda, a.wbf + a.ld.sig, D, C_s, ldc, 2 * R, D);
  tile_mma<TM, false, false>(A_s, lda, a.wbf + a.ld.gate, D, C_s + D, ldc, 2 * R, D);
  __syncthreads();
  for (int idx = threadIdx.x; idx < TM * D; idx += NT) {
    const int r = idx / D, d = idx % D, t = t0 + r;
    float vs = C_s[r * ldc + d], vg = C_s[r * ldc + D + d];
    if (a.ld.sig_b >= 0) {
      vs += a.params[a.ld.sig_b + d];
      vg += a.params[a.ld.gate_b + d];
    }
    if (a.gc_tbl != nullptr && t < T) {
      int id = a.ids[(size_t)b * T + t];
      id = min(max(id, 0), a.C1 - 1);
      const float* g = a.gc_tbl + (size_t)id * 2 * D;
      vs += g[d];
      vg += g[D + d];
    }
    const bf16 zz = f2bf(tanh_fast(vs) * sigmoid_fast(vg));
    Z_s[r * ldz + d] = zz;
    if (t < T) a.z[((size_t)b * T + t) * a.LD + a.l * D + d] = zz;
  }
  if (a.xout == nullptr) return;
  __syncthreads();
  tile_mma<TM, false, false>(Z_s, ldz, a.wbf + a.ld.res, R, C2_s, ldc2, D, R);
  __syncthreads();
  for (int idx = threadIdx.x; idx < TM * R; idx += NT) {
    const int r = idx / R, c = idx % R, t = t0 + r;
    if (t < T) {
      float v = bf2f(A_s[r * lda + R + c]) + C2_s[r * ldc2 + c];
      if (a.ld.res_b >= 0) v += a.params[a.ld.res_b + c];
      a.xout[((size_t)b * (a.dil_next + T) + a.dil_next + t) * R + c] = f2bf(v);
    }
  }
}

static size_t layer_fwd_smem(int R, int D) {
  return (size_t)TM * (2 * R + 8) * 2 + (size_t)TM * (2 * D + 4) * 4 + (size_t)TM * (D + 8) * 2 +
         (size_t)TM * (R + 4) * 4;
}

// ======================================================================================
// skip GEMM (concat-K over layers) + ReLU + POST1 + ReLU + POST2 + masked softmax-xent
// (tmodel.py:321-324, 187-215, 228-249); emits dlogits = (softmax - onehot)*mask (bf16)
// ======================================================================================
struct PostArgs {
  const bf16* wbf;
  const float* params;
  const LayerDesc* layers;
  int64_t off_post1, off_post1_b, off_post2, off_post2_b;
  const float* skip_bias;
  const bf16* z;
  bf16* h1;
  bf16* h2;
  bf16* dlogits;
  bf16* dp1;
  bf16* dskip;
  bf16* dz;
  float* logits_out;
  float* grads;  // backward only
  const int32_t* wav;
  const int32_t* ids;
  double* stats;
  int B, T, D, S, P, Q, L, LD, use_bias;
  int64_t rows;
  int AW, CW;  // shared tile widths
};

__global__ void __launch_bounds__(NT) k_post_fwd(PostArgs a) {
  using namespace nvcuda;
  extern __shared__ __align__(128) unsigned char smem[];
  const int lda = a.AW + 8, ldc = a.CW + 4;
  bf16* A_s = reinterpret_cast<bf16*>(smem);  // also H_s
  float* C_s = reinterpret_cast<float*>(A_s + TM * lda);
  __shared__ double red[3][NT / 32];
  const int D = a.D, S = a.S, P = a.P, Q = a.Q, T = a.T;
  const int64_t row0 = (int64_t)blockIdx.x * TM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int lpc = a.AW / D;  // layers per K chunk

  for (int l0 = 0; l0 < a.L; l0 += lpc) {
    const int nl = min(lpc, a.L - l0), kc = nl * D, cpr = kc / 8;
    __syncthreads();
    for (int idx = threadIdx.x; idx < TM * cpr; idx += NT) {
      const int r = idx / cpr, c = idx % cpr;
      const int64_t row = row0 + r;
      copy16(A_s + r * lda + c * 8, a.z + (size_t)row * a.LD + l0 * D + c * 8, row < a.rows);
    }
    __syncthreads();
    const int ntn = S >> 4, ntiles = (TM / 16) * ntn;
    for (int tile = warp; tile < ntiles; tile += NT / 32) {
      const int mi = tile % (TM / 16), ni = tile / (TM / 16);
      wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc;
      float* cptr = C_s + mi * 16 * ldc + ni * 16;
      if (l0 > 0)
        wmma::load_matrix_sync(acc, cptr, ldc, wmma::mem_row_major);
      else
        wmma::fill_fragment(acc, 0.f);
      for (int j = 0; j < nl; ++j) {
        const bf16* Bg = a.wbf + a.layers[l0 + j].skip;
        for (int kk = 0; kk < D; kk += 16) {
          wmma::fragment<wmma::matrix_a, 16, 16, 16, bf16, wmma::row_major> fa;
          wmma::fragment<wmma::matrix_b, 16, 16, 16, bf16, wmma::row_major> fb;
          wmma::load_matrix_sync(fa, A_s + mi * 16 * lda + j * D + kk, lda);
          wmma::load_matrix_sync(fb, Bg + (size_t)kk * S + ni * 16, S);
          wmma::mma_sync(acc, fa, fb, acc);
        }
      }
      wmma::store_matrix_sync(cptr, acc, ldc, wmma::mem_row_major);
    }
  }
  __syncthreads();
  // h1 = relu(skip_sum + sum_l bias_l)
  for (int idx = threadIdx.x; idx < TM * S; idx += NT) {
    const int r = idx / S, s = idx % S;
    float v = C_s[r * ldc + s];
    if (a.use_bias) v += a.skip_bias[s];
    const bf16 h = f2bf(fmaxf(v, 0.f));
    A_s[r * lda + s] = h;
    if (row0 + r < a.rows) a.h1[(size_t)(row0 + r) * S + s] = h;
  }
  __syncthreads();
  tile_mma<TM, false, false>(A_s, lda, a.wbf + a.off_post1, P, C_s, ldc, S, P);
  __syncthreads();
  for (int idx = threadIdx.x; idx < TM * P; idx += NT) {
    const int r = idx / P, p = idx % P;
    float v = C_s[r * ldc + p];
    if (a.use_bias) v += a.params[a.off_post1_b + p];
    const bf16 h = f2bf(fmaxf(v, 0.f));
    A_s[r * lda + p] = h;
    if (row0 + r < a.rows) a.h2[(size_t)(row0 + r) * P + p] = h;
  }
  __syncthreads();
  tile_mma<TM, false, false>(A_s, lda, a.wbf + a.off_post2, Q, C_s, ldc, P, Q);
  __syncthreads();
  // masked softmax cross entropy, one warp per row, lane owns 8 consecutive logits (Q == 256)
  float acc_x = 0.f, acc_n = 0.f, acc_d = 0.f;
  for (int r = warp; r < TM; r += NT / 32) {
    const int64_t row = row0 + r;
    if (row >= a.rows) continue;
    const int b = (int)(row / T), t = (int)(row % T);
    float v[8];
    float mx = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = C_s[r * ldc + lane * 8 + j] + (a.use_bias ? a.params[a.off_post2_b + lane * 8 + j] : 0.f);
      if (v[j] > mx) {
        mx = v[j];
        arg = lane * 8 + j;
      }
    }
    if (a.logits_out != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a.logits_out[(size_t)row * Q + lane * 8 + j] = v[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {  // argmax, smallest index on ties (tf.argmax)
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (om > mx || (om == mx && oa < arg)) {
        mx = om;
        arg = oa;
      }
    }
    float e[8], sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      e[j] = __expf(v[j] - mx);
      sum += e[j];
    }
    sum = warp_sum(sum);
    const bool valid = (t + 1 < T) && (a.ids[(size_t)b * T + t + 1] != 0);  // tmodel.py:232
    const int label_raw = (t + 1 < T) ? a.wav[(size_t)b * T + t + 1] : 0;    // tmodel.py:230
    const bool label_ok = label_raw >= 0 && label_raw < Q;                   // out of range: all-zero one-hot row (tmodel.py:64)
    const int label = label_ok ? label_raw : -1;
    const int lsel = label_ok ? label : 0;
    const float vl = __shfl_sync(0xffffffffu, v[lsel & 7], lsel >> 3);
    // NB: v[label&7] with a lane-varying index would be needed if label differed per lane; it is
    // warp-uniform here, so every lane evaluates the same register select.
    const float xent = label_ok ? __logf(sum) + mx - vl : 0.f;
    const float inv = 1.f / sum;
    __align__(16) bf16 dl[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float g = e[j] * inv - ((lane * 8 + j) == label ? 1.f : 0.f);
      dl[j] = f2bf(valid ? g : 0.f);
    }
    *reinterpret_cast<uint4*>(a.dlogits + (size_t)row * Q + lane * 8) = *reinterpret_cast<uint4*>(dl);
    if (valid) {
      acc_x += xent;
      acc_n += 1.f;
      acc_d += fabsf((float)(lsel - arg));
    }
  }
  if (lane == 0) {
    red[0][warp] = acc_x;
    red[1][warp] = acc_n;
    red[2][warp] = acc_d;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0;
    for (int w = 0; w < NT / 32; ++w) s += red[threadIdx.x][w];
    if (s != 0.0) atomicAdd(a.stats + threadIdx.x, s);
  }
}

// ======================================================================================
// post-net backward: dlogits -> dp1 -> dskip -> dz(skip part) for every layer; bias grads
// ======================================================================================
__global__ void __launch_bounds__(NT) k_post_bwd(PostArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lda = a.AW + 8, ldc = a.CW + 4;
  bf16* A_s = reinterpret_cast<bf16*>(smem);
  float* C_s = reinterpret_cast<float*>(A_s + TM * lda);
  const int D = a.D, S = a.S, P = a.P, Q = a.Q;
  const int64_t row0 = (int64_t)blockIdx.x * TM;
  // dlogits tile
  for (int idx = threadIdx.x; idx < TM * (Q / 8); idx += NT) {
    const int r = idx / (Q / 8), c = idx % (Q / 8);
    copy16(A_s + r * lda + c * 8, a.dlogits + (size_t)(row0 + r) * Q + c * 8, row0 + r < a.rows);
  }
  __syncthreads();
  if (a.use_bias) {  // POST2_BIAS grad = colsum(dlogits)
    for (int q = threadIdx.x; q < Q; q += NT) {
      float s = 0.f;
      for (int r = 0; r < TM; ++r) s += bf2f(A_s[r * lda + q]);
      if (s != 0.f) atomicAdd(a.grads + a.off_post2_b + q, s);
    }
  }
  // dp1 = (dlogits @ POST2^T) * (h2 > 0)
  tile_mma<TM, true, false>(A_s, lda, a.wbf + a.off_post2, Q, C_s, ldc, Q, P);
  __syncthreads();
  for (int idx = threadIdx.x; idx < TM * P; idx += NT) {
    const int r = idx / P, p = idx % P;
    const int64_t row = row0 + r;
    bf16 g = f2bf(0.f);
    if (row < a.rows) {
      const float h = bf2f(a.h2[(size_t)row * P + p]);
      g = f2bf(h > 0.f ? C_s[r * ldc + p] : 0.f);
      a.dp1[(size_t)row * P + p] = g;
    }
    A_s[r * lda + p] = g;
  }
  __syncthreads();
  if (a.use_bias) {
    for (int p = threadIdx.x; p < P; p += NT) {
      float s = 0.f;
      for (int r = 0; r < TM; ++r) s += bf2f(A_s[r * lda + p]);
      if (s != 0.f) atomicAdd(a.grads + a.off_post1_b + p, s);
    }
  }
  // dskip = (dp1 @ POST1^T) * (h1 > 0)
  tile_mma<TM, true, false>(A_s, lda, a.wbf + a.off_post1, P, C_s, ldc, P, S);
  __syncthreads();
  for (int idx = threadIdx.x; idx < TM * S; idx += NT) {
    const int r = idx / S, s = idx % S;
    const int64_t row = row0 + r;
    bf16 g = f2bf(0.f);
    if (row < a.rows) {
      const float h = bf2f(a.h1[(size_t)row * S + s]);
      g = f2bf(h > 0.f ? C_s[r * ldc + s] : 0.f);
      a.dskip[(size_t)row * S + s] = g;
    }
    A_s[r * lda + s] = g;
  }
  __syncthreads();
  if (a.use_bias) {  // every SKIP_BIAS_l receives colsum(dskip); accumulate once into layer 0's
    for (int s = threadIdx.x; s < S; s += NT) {  // slot, k_bcast_skip_bias copies it afterwards
      float sum = 0.f;
      for (int r = 0; r < TM; ++r) sum += bf2f(A_s[r * lda + s]);
      if (sum != 0.f) atomicAdd(a.grads + a.layers[0].skip_b + s, sum);
    }
  }
  // dz_l(skip part) = dskip @ SKIP_l^T for every layer, CW/D layers per pass
  const int lpp = a.CW / D;
  for (int l0 = 0; l0 < a.L; l0 += lpp) {
    const int nl = min(lpp, a.L - l0);
    __syncthreads();
    for (int j = 0; j < nl; ++j)
      tile_mma<TM, true, false>(A_s, lda, a.wbf + a.layers[l0 + j].skip, S, C_s + j * D, ldc, S, D);
    __syncthreads();
    for (int idx = threadIdx.x; idx < TM * nl * D; idx += NT) {
      const int r = idx / (nl * D), c = idx % (nl * D);
      const int64_t row = row0 + r;
      if (row < a.rows) a.dz[((size_t)(l0 + c / D) * a.rows + row) * D + c % D] = f2bf(C_s[r * ldc + c]);  // [L][B*T][D]
    }
  }
}

__global__ void k_bcast_skip_bias(float* grads, const LayerDesc* layers, int L, int S) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const float v = grads[layers[0].skip_b + s];
  for (int l = 1; l < L; ++l) grads[layers[l].skip_b + s] = v;
}

static size_t post_smem(int AW, int CW) { return (size_t)TM * (AW + 8) * 2 + (size_t)TM * (CW + 4) * 4; }

// ======================================================================================
// layer backward, part A: recompute the gate pre-activations, dz = dz_skip + dx' @ Wr^T,
// dv = gate'(v) * dz  (stored bf16 [B*T][2D]); bias / GC-table gradients
// ======================================================================================
struct LayerBwdArgs {
  const bf16* wbf;
  const float* params;
  LayerDesc ld;
  const bf16* xin;      // xfull_l
  const bf16* dx_next;  // [B*T][R] gradient wrt x_{l+1}, nullptr for the last layer
  bf16* dx_out;         // [B*T][R] gradient wrt x_l
  const bf16* dz;       // [L][B*T][D]
  bf16* dv;             // [B*T][2D]
  const float* gc_tbl;
  float* dgc_tbl;       // this layer's [C1][2D] gradient table
  const int32_t* ids;
  float* grads;
  int B, T, R, D, LD, l, C1;
};

__global__ void __launch_bounds__(NT) k_layer_bwd_a(LayerBwdArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int R = a.R, D = a.D;
  const int lda = 2 * R + 8, ldc = 2 * D + 4, ldx = R + 8, ldc2 = D + 4, ldv = 2 * D + 8;
  bf16* A_s = reinterpret_cast<bf16*>(smem);
  float* C_s = reinterpret_cast<float*>(A_s + TM * lda);
  bf16* X_s = reinterpret_cast<bf16*>(C_s + TM * ldc);
  float* C2_s = reinterpret_cast<float*>(X_s + TM * ldx);
  bf16* V_s = reinterpret_cast<bf16*>(C2_s + TM * ldc2);
  __shared__ int id_s[TM];
  const int b = blockIdx.y, t0 = blockIdx.x * TM, T = a.T, dil = a.ld.dil;
  const bf16* xb = a.xin + (size_t)b * (dil + T) * R;
  load_conv_tile(A_s, lda, xb, t0, T, dil, R);
  if (a.dx_next != nullptr) {
    for (int idx = threadIdx.x; idx < TM * (R / 8); idx += NT) {
      const int r = idx / (R / 8), c = idx % (R / 8), t = t0 + r;
      copy16(X_s + r * ldx + c * 8, a.dx_next + ((size_t)b * T + t) * R + c * 8, t < T);
    }
  }
  if (threadIdx.x < TM) {
    const int t = t0 + threadIdx.x;
    int id = (t < T) ? a.ids[(size_t)b * T + t] : -1;
    id_s[threadIdx.x] = id < 0 ? -1 : min(id, a.C1 - 1);
  }
  __syncthreads();
  tile_mma<TM, false, false>(A_s, lda, a.wbf + a.ld.sig, D, C_s, ldc, 2 * R, D);
  tile_mma<TM, false, false>(A_s, lda, a.wbf + a.ld.gate, D, C_s + D, ldc, 2 * R, D);
  if (a.dx_next != nullptr)  // dz(res part) = dx' @ RESIDUAL^T  (RESIDUAL stored [D][R])
    tile_mma<TM, true, false>(X_s, ldx, a.wbf + a.ld.res, R, C2_s, ldc2, R, D);
  __syncthreads();
  for (int idx = threadIdx.x; idx < TM * D; idx += NT) {
    const int r = idx / D, d = idx % D, t = t0 + r;
    bf16 o_s = f2bf(0.f), o_g = f2bf(0.f);
    if (t < T) {
      float vs = C_s[r * ldc + d], vg = C_s[r * ldc + D + d];
      if (a.ld.sig_b >= 0) {
        vs += a.params[a.ld.sig_b + d];
        vg += a.params[a.ld.gate_b + d];
      }
      if (a.gc_tbl != nullptr) {
        const float* g = a.gc_tbl + (size_t)max(id_s[r], 0) * 2 * D;
        vs += g[d];
        vg += g[D + d];
      }
      const float th = tanh_fast(vs), sg = sigmoid_fast(vg);
      float dz = bf2f(a.dz[(((size_t)a.l * a.B + b) * T + t) * D + d]);
      if (a.dx_next != nullptr) dz += C2_s[r * ldc2 + d];
      o_s = f2bf(dz * sg * (1.f - th * th));
      o_g = f2bf(dz * th * sg * (1.f - sg));
      a.dv[((size_t)b * T + t) * 2 * D + d] = o_s;
      a.dv[((size_t)b * T + t) * 2 * D + D + d] = o_g;
    }
    V_s[r * ldv + d] = o_s;
    V_s[r * ldv + D + d] = o_g;
  }
  __syncthreads();
  // column sums of dv: SIGNAL_BIAS / GATE_BIAS gradients, and the GC-table gradient segmented by id
  for (int n = threadIdx.x; n < 2 * D; n += NT) {
    float tot = 0.f, seg = 0.f;
    int cur = -2;
    for (int r = 0; r < TM; ++r) {
      const float v = bf2f(V_s[r * ldv + n]);
      tot += v;
      if (a.dgc_tbl != nullptr) {
        const int id = id_s[r];
        if (id != cur) {
          if (cur >= 0 && seg != 0.f) atomicAdd(a.dgc_tbl + (size_t)cur * 2 * D + n, seg);
          cur = id;
          seg = 0.f;
        }
        seg += v;
      }
    }
    if (a.dgc_tbl != nullptr && cur >= 0 && seg != 0.f) atomicAdd(a.dgc_tbl + (size_t)cur * 2 * D + n, seg);
    if (a.ld.sig_b >= 0 && tot != 0.f)
      atomicAdd(a.grads + (n < D ? a.ld.sig_b + n : a.ld.gate_b + (n - D)), tot);
  }
  if (a.dx_next != nullptr && a.ld.res_b >= 0) {  // RESIDUAL_BIAS grad = colsum(dx')
    for (int c = threadIdx.x; c < R; c += NT) {
      float s = 0.f;
      for (int r = 0; r < TM; ++r) s += bf2f(X_s[r * ldx + c]);
      if (s != 0.f) atomicAdd(a.grads + a.ld.res_b + c, s);
    }
  }
}

static size_t layer_bwd_a_smem(int R, int D) {
  return (size_t)TM * (2 * R + 8) * 2 + (size_t)TM * (2 * D + 4) * 4 + (size_t)TM * (R + 8) * 2 +
         (size_t)TM * (D + 4) * 4 + (size_t)TM * (2 * D + 8) * 2;
}

// part B: dx_l[t] = dx_{l+1}[t] + dv[t] @ W[1]^T + dv[t+dil] @ W[0]^T   (rows t+dil >= T drop out:
// the SAVE prefix is a variable, not a graph tensor -- gradients stop at the stage boundary)
__global__ void __launch_bounds__(NT) k_layer_bwd_b(LayerBwdArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int R = a.R, D = a.D;
  const int lda = 4 * D + 8, ldc = R + 4;
  bf16* A_s = reinterpret_cast<bf16*>(smem);
  float* C_s = reinterpret_cast<float*>(A_s + TM * lda);
  const int b = blockIdx.y, t0 = blockIdx.x * TM, T = a.T, dil = a.ld.dil;
  const int cpr = 4 * D / 8, half_c = 2 * D / 8;
  for (int idx = threadIdx.x; idx < TM * cpr; idx += NT) {
    const int r = idx / cpr, c = idx % cpr;
    const int half = c >= half_c;
    const int cc = c - half * half_c;
    const int t = t0 + r + (half ? dil : 0);
    copy16(A_s + r * lda + c * 8, a.dv + ((size_t)b * T + t) * 2 * D + cc * 8, t < T);
  }
  __syncthreads();
  // B[k=d][n=r] = W[tap][r][d]  -> stored [R][D] row-major == col-major B with ldb = D
  const bf16* Wsig = a.wbf + a.ld.sig;
  const bf16* Wgate = a.wbf + a.ld.gate;
  const size_t tap = (size_t)R * D;
  tile_mma<TM, true, false>(A_s, lda, Wsig + tap, D, C_s, ldc, D, R);            // dvs[t]   x SIGNAL[1]
  __syncthreads();
  tile_mma<TM, true, true>(A_s + D, lda, Wgate + tap, D, C_s, ldc, D, R);        // dvg[t]   x GATE[1]
  __syncthreads();
  tile_mma<TM, true, true>(A_s + 2 * D, lda, Wsig, D, C_s, ldc, D, R);           // dvs[t+d] x SIGNAL[0]
  __syncthreads();
  tile_mma<TM, true, true>(A_s + 3 * D, lda, Wgate, D, C_s, ldc, D, R);          // dvg[t+d] x GATE[0]
  __syncthreads();
  for (int idx = threadIdx.x; idx < TM * R; idx += NT) {
    const int r = idx / R, c = idx % R, t = t0 + r;
    if (t < T) {
      float v = C_s[r * ldc + c];
      if (a.dx_next != nullptr) v += bf2f(a.dx_next[((size_t)b * T + t) * R + c]);
      a.dx_out[((size_t)b * T + t) * R + c] = f2bf(v);
    }
  }
}

static size_t layer_bwd_b_smem(int R, int D) {
  return (size_t)TM * (4 * D + 8) * 2 + (size_t)TM * (R + 4) * 4;
}

// ======================================================================================
// generic weight gradient: out[ka][n] += sum_{b,t} A[b][t+a_row_off][a_col0+ka] * Y[b][t][y_col0+n]
// ======================================================================================
struct WgradArgs {
  const bf16* A;
  int64_t a_slot_pitch;
  int a_row_off, lda, a_col0;
  const bf16* Y;
  int64_t y_slot_pitch;
  int ldy, y_col0;
  float* out;
  int ldo, Ka, N, B, T;
  int64_t rows_per_cta;
};

template <int NF>
__global__ void __launch_bounds__(NT) k_wgrad(WgradArgs a) {
  using namespace nvcuda;
  constexpr int KA_BLK = 32, N_BLK = 64 * NF, LDA = KA_BLK + 8, LDY = N_BLK + 8;
  __shared__ __align__(128) bf16 A_c[TM * LDA];
  __shared__ __align__(128) bf16 Y_c[TM * LDY];
  __shared__ __align__(128) float stage[NT / 32][16 * 16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ka0 = blockIdx.y * KA_BLK, n0 = blockIdx.z * N_BLK;
  const int64_t rows = (int64_t)a.B * a.T;
  const int64_t r_begin = (int64_t)blockIdx.x * a.rows_per_cta;
  const int64_t r_end = min(rows, r_begin + a.rows_per_cta);
  const int kt = warp & 1, nq = warp >> 1;  // warp owns ka tile kt and n tiles nq + 4*j
  wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[NF];
#pragma unroll
  for (int j = 0; j < NF; ++j) wmma::fill_fragment(acc[j], 0.f);
  for (int64_t rc = r_begin; rc < r_end; rc += TM) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < TM * (KA_BLK / 8); idx += NT) {
      const int r = idx / (KA_BLK / 8), c = idx % (KA_BLK / 8);
      const int64_t row = rc + r;
      const int64_t b = row / a.T, t = row % a.T;
      const bool ok = row < r_end && (ka0 + c * 8) < a.Ka;
      copy16(A_c + r * LDA + c * 8,
             a.A + b * a.a_slot_pitch + (t + a.a_row_off) * (int64_t)a.lda + a.a_col0 + ka0 + c * 8, ok);
    }
    for (int idx = threadIdx.x; idx < TM * (N_BLK / 8); idx += NT) {
      const int r = idx / (N_BLK / 8), c = idx % (N_BLK / 8);
      const int64_t row = rc + r;
      const int64_t b = row / a.T, t = row % a.T;
      const bool ok = row < r_end && (n0 + c * 8) < a.N;
      copy16(Y_c + r * LDY + c * 8, a.Y + b * a.y_slot_pitch + t * (int64_t)a.ldy + a.y_col0 + n0 + c * 8, ok);
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < TM / 16; ++ks) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, bf16, wmma::col_major> fa;
      wmma::load_matrix_sync(fa, A_c + ks * 16 * LDA + kt * 16, LDA);
#pragma unroll
      for (int j = 0; j < NF; ++j) {
        wmma::fragment<wmma::matrix_b, 16, 16, 16, bf16, wmma::row_major> fb;
        wmma::load_matrix_sync(fb, Y_c + ks * 16 * LDY + (nq + 4 * j) * 16, LDY);
        wmma::mma_sync(acc[j], fa, fb, acc[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NF; ++j) {
    const int ka = ka0 + kt * 16, n = n0 + (nq + 4 * j) * 16;
    if (ka < a.Ka && n < a.N) {  // warp-uniform
      wmma::store_matrix_sync(stage[warp], acc[j], 16, wmma::mem_row_major);
      __syncwarp();
      for (int e = lane; e < 256; e += 32) {
        const float v = stage[warp][e];
        if (v != 0.f) atomicAdd(a.out + (size_t)(ka + e / 16) * a.ldo + n + e % 16, v);
      }
      __syncwarp();
    }
  }
}

// ======================================================================================
// PRE gather backward (scatter-add) + PRE_BIAS
// ======================================================================================
// dx0 = gradient wrt the layer-0 input; with p0 != nullptr it arrives in split form dx0[t] = dx0[t] + p0[t + dil0]
// (rows t + dil0 >= T contribute nothing: truncated at the stage boundary), see layer_umma.cu.
// One warp per row, lane <-> channel: the shared-memory atomics of a warp hit 32 different banks.  Rows whose code is
// out of range (an all-zero one-hot row, tmodel.py:64) still feed PRE_BIAS: they go to table row Q, and PRE_BIAS's
// gradient is the column sum of the whole table (no per-row atomic on 32 hot addresses).
__global__ void __launch_bounds__(1024) k_embed_bwd(const bf16* __restrict__ dx0, const bf16* __restrict__ p0, int dil0,
                                                    int T, const int32_t* __restrict__ wav, float* __restrict__ part,
                                                    int64_t rows, int R, int Q, int64_t rows_per_cta, int ncopy) {
  extern __shared__ float tbl[];  // [ncopy][Q + 1][R]
  const int nt = blockDim.x;
  for (int i = threadIdx.x; i < ncopy * (Q + 1) * R; i += nt) tbl[i] = 0.f;
  __syncthreads();
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  if (R == 32) {
    // audio codes cluster around mid-scale, so many warps hit the same table rows: NCOPY private copies of the
    // table (warp w uses copy w % NCOPY) cut the same-address serialisation of the shared-memory atomics
    // A lane owns 8 channels (16 bytes) of a row: a warp covers 8 rows per load instruction and keeps UNR of them in
    // flight -- 2 x UNR x 512 bytes per warp.  (One 2-byte element per lane, as before, left ~8 KB in flight per SM and
    // the kernel bound by DRAM latency: 93 us for 64 MB.)
    constexpr int UNR = 4;
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, nw = nt / 32;
    const int sub = lane >> 2, ch = lane & 3;  // row within the group of 8, 16-byte chunk of the row
    float* mine = tbl + (size_t)(wrp % ncopy) * (Q + 1) * 32;
    for (int64_t g0 = r0 + (int64_t)wrp * 8 * UNR; g0 < r1; g0 += (int64_t)nw * 8 * UNR) {
      uint4 a[UNR], b[UNR];
      int code[UNR];
      bool ok[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int64_t row = g0 + u * 8 + sub;
        a[u] = b[u] = make_uint4(0u, 0u, 0u, 0u);
        code[u] = 0;
        ok[u] = row < r1;
        if (ok[u]) {
          a[u] = __ldg(reinterpret_cast<const uint4*>(dx0 + row * 32) + ch);
          if (p0 != nullptr && (int)(row % T) + dil0 < T) b[u] = __ldg(reinterpret_cast<const uint4*>(p0 + (row + dil0) * 32) + ch);
          code[u] = __ldg(wav + row);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (!ok[u]) continue;
        const int c = (code[u] < 0 || code[u] >= Q) ? Q : code[u];
        float* dst = mine + c * 32 + ch * 8;
        const uint32_t wa[4] = {a[u].x, a[u].y, a[u].z, a[u].w}, wb[4] = {b[u].x, b[u].y, b[u].z, b[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          atomicAdd(dst + 2 * j, __uint_as_float(wa[j] << 16) + __uint_as_float(wb[j] << 16));
          atomicAdd(dst + 2 * j + 1, __uint_as_float(wa[j] & 0xffff0000u) + __uint_as_float(wb[j] & 0xffff0000u));
        }
      }
    }
  } else {
    for (int64_t i = r0 * R + threadIdx.x; i < r1 * R; i += nt) {
      const int64_t row = i / R;
      const int r = (int)(i % R);
      float v = bf2f(dx0[i]);
      if (p0 != nullptr && (int)(row % T) + dil0 < T) v += bf2f(p0[i + (int64_t)dil0 * R]);
      int code = wav[row];
      if (code < 0 || code >= Q) code = Q;
      atomicAdd(&tbl[code * R + r], v);
    }
  }
  __syncthreads();
  float* dst = part + (size_t)blockIdx.x * (Q + 1) * R;  // per-CTA partial table: no global atomics
  for (int i = threadIdx.x; i < (Q + 1) * R; i += nt) {
    float sum = 0.f;
    for (int c = 0; c < ncopy; ++c) sum += tbl[(size_t)c * (Q + 1) * R + i];
    dst[i] = sum;
  }
}
// grads[PRE] += sum over the partial tables (rows < Q); grads[PRE_BIAS] += column sums over all Q + 1 rows
__global__ void k_embed_bwd_reduce(const float* __restrict__ part, int n_part, float* grads, int64_t off_pre,
                                   int64_t off_pre_b, int R, int Q) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (Q + 1) * R) return;
  float sum = 0.f;
  for (int p = 0; p < n_part; ++p) sum += part[(size_t)p * (Q + 1) * R + i];
  if (i < Q * R) grads[off_pre + i] += sum;
  if (off_pre_b >= 0 && sum != 0.f) atomicAdd(grads + off_pre_b + i % R, sum);
}

// ======================================================================================
// GC backward from the per-layer table gradients
// ======================================================================================
// dGC_EMBED[c][g] = sum_l sum_n dTbl[l][c][n] * Wgc_l[g][n]
__global__ void k_gc_bwd_embed(const float* __restrict__ p, const float* __restrict__ dtbl,
                               const LayerDesc* __restrict__ layers, float* grads, int64_t off_embed,
                               int L, int C1, int G, int D) {
  const int c = blockIdx.x, g = threadIdx.x;
  if (g >= G) return;
  float acc = 0.f;
  for (int l = 0; l < L; ++l) {
    const float* dt = dtbl + ((size_t)l * C1 + c) * 2 * D;
    const float* ws = p + layers[l].gc_sig + (size_t)g * D;
    const float* wg = p + layers[l].gc_gate + (size_t)g * D;
    for (int n = 0; n < D; ++n) acc += dt[n] * ws[n] + dt[D + n] * wg[n];
  }
  grads[off_embed + (size_t)c * G + g] = acc;
}
// dGC_{SIGNAL,GATE}_l[g][n] = sum_c GC_EMBED[c][g] * dTbl[l][c][n]
__global__ void k_gc_bwd_proj(const float* __restrict__ p, const float* __restrict__ dtbl,
                              const LayerDesc* __restrict__ layers, float* grads, int64_t off_embed,
                              int C1, int G, int D) {
  const int l = blockIdx.y, g = blockIdx.x, n = threadIdx.x;
  if (n >= 2 * D) return;
  float acc = 0.f;
  for (int c = 0; c < C1; ++c) acc += p[off_embed + (size_t)c * G + g] * dtbl[((size_t)l * C1 + c) * 2 * D + n];
  if (n < D)
    grads[layers[l].gc_sig + (size_t)g * D + n] = acc;
  else
    grads[layers[l].gc_gate + (size_t)g * D + (n - D)] = acc;
}

// ======================================================================================
// optimiser + L2
// ======================================================================================
__global__ void k_adam(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, const uint8_t* __restrict__ kind,
                       const double* __restrict__ n_valid, int64_t n, float lr_t, float l2_factor,
                       float beta1, float beta2, float eps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double nv = *n_valid;
  const float scale = nv > 0.0 ? (float)(1.0 / nv) : 0.f;  // tmodel.py:246-249
  const float wi = w[i];
  float gi = g[i] * scale;
  if (kind[i]) gi += l2_factor * wi;  // d/dw of l2_factor * 0.5*|w|^2, filters only (tmodel.py:252-261)
  const float mi = beta1 * m[i] + (1.f - beta1) * gi;
  const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  w[i] = wi - lr_t * mi / (sqrtf(vi) + eps);
}

__global__ void k_l2(const float* __restrict__ w, const uint8_t* __restrict__ kind, int64_t n, double* out) {
  __shared__ double red[NT / 32];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (kind[i]) acc += 0.5 * (double)w[i] * (double)w[i];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < NT / 32; ++i) s += red[i];
    atomicAdd(out, s);
  }
}

// `add` (optional, same [rows][ld] geometry): dst += add[t + add_shift] for t + add_shift < slot_rows
// loader: mu-law codes travel host -> device as uint8 (5 bytes per timestep with the int32 id) and are
// widened here to the int32 the kernels index with (data.py:262-265 dtypes).  16 codes per thread, grid-stride.
__global__ void k_widen_u8(const uint8_t* __restrict__ src, int32_t* __restrict__ dst, int64_t n) {
  const int64_t n16 = n / 16;
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n + 15) / 16; i += (int64_t)gridDim.x * blockDim.x) {
    if (aligned && i < n16) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(src) + i);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      int4* o = reinterpret_cast<int4*>(dst) + 4 * i;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        o[k] = make_int4(w[k] & 0xff, (w[k] >> 8) & 0xff, (w[k] >> 16) & 0xff, w[k] >> 24);
    } else {
      for (int64_t j = 16 * i; j < n && j < 16 * i + 16; ++j) dst[j] = src[j];
    }
  }
}

__global__ void k_debug_read(const bf16* __restrict__ src, float* __restrict__ dst, int64_t n_rows, int ncols,
                             int64_t slot_rows, int64_t slot_pitch_rows, int row_off, int ld, int col0,
                             const bf16* __restrict__ add, int add_shift) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * ncols) return;
  const int64_t row = i / ncols;
  const int c = (int)(i % ncols);
  const int64_t b = row / slot_rows, t = row % slot_rows;
  float v = bf2f(src[(b * slot_pitch_rows + t + row_off) * ld + col0 + c]);
  if (add != nullptr && t + add_shift < slot_rows) v += bf2f(add[(b * slot_pitch_rows + t + add_shift) * ld + col0 + c]);
  dst[i] = v;
}

// ======================================================================================
// host-side orchestration
// ======================================================================================
int ensure_layer_table(wn_model* m) {
  if (m->sm_count == 0) {
    int dev = 0;
    WN_CUDA_CHECK(cudaGetDevice(&dev));
    WN_CUDA_CHECK(cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  if (m->d_layers == nullptr) {
    WN_CUDA_CHECK(cudaMalloc(&m->d_layers, sizeof(LayerDesc) * m->L));
    WN_CUDA_CHECK(cudaMemcpy(m->d_layers, m->layers.data(), sizeof(LayerDesc) * m->L, cudaMemcpyHostToDevice));
    m->d_layers_T = -1;
  }
  return WN_OK;
}

static int ensure_kind(wn_model* m) {
  if (m->d_kind == nullptr) {
    std::vector<uint8_t> kind((size_t)m->n_param_elems, 0);
    for (const ParamEntry& e : m->params)
      if (e.kind == WN_KIND_FILTER)
        for (int64_t i = 0; i < e.numel(); ++i) kind[(size_t)(e.offset + i)] = 1;
    WN_CUDA_CHECK(cudaMalloc(&m->d_kind, kind.size()));
    WN_CUDA_CHECK(cudaMemcpy(m->d_kind, kind.data(), kind.size(), cudaMemcpyHostToDevice));
  }
  return WN_OK;
}

static int ensure_tables(wn_model* m, int T) {
  int rc = ensure_layer_table(m);
  if (rc) return rc;
  const WorkspaceLayout& wl = workspace_layout(m, T);
  if (m->d_layers_T != T) {
    // slice_sz changed: the per-layer workspace offsets move.  Rare; drain the device first so no
    // in-flight kernel still reads the old table.
    WN_CUDA_CHECK(cudaDeviceSynchronize());
    for (int l = 0; l < m->L; ++l) m->layers[l].xfull_off = wl.xfull[l];
    WN_CUDA_CHECK(cudaMemcpy(m->d_layers, m->layers.data(), sizeof(LayerDesc) * m->L, cudaMemcpyHostToDevice));
    m->d_layers_T = T;
  }
  return ensure_kind(m);
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    WN_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return WN_OK;
}

static Dims make_dims(const wn_model* m, int T) {
  Dims d;
  d.B = m->n_slots; d.T = T; d.R = m->a.n_res; d.D = m->a.n_dil; d.S = m->a.n_skip; d.P = m->a.n_post;
  d.Q = m->a.n_quant; d.L = m->L; d.LD = m->L * m->a.n_dil; d.G = m->a.n_gc_embed;
  d.C1 = m->a.n_gc_category + 1; d.use_bias = m->a.use_bias; d.rows = (int64_t)m->n_slots * T;
  return d;
}

static PostArgs make_post_args(wn_model* m, const Dims& d, const WorkspaceLayout& wl, unsigned char* ws,
                               const float* params) {
  PostArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.wbf = reinterpret_cast<const bf16*>(ws + wl.wbf);
  pa.params = params;
  pa.layers = m->d_layers;
  pa.off_post1 = m->off_post1; pa.off_post1_b = m->off_post1_b;
  pa.off_post2 = m->off_post2; pa.off_post2_b = m->off_post2_b;
  pa.skip_bias = reinterpret_cast<const float*>(ws + wl.skip_bias);
  pa.z = reinterpret_cast<const bf16*>(ws + wl.z);
  pa.h1 = reinterpret_cast<bf16*>(ws + wl.h1);
  pa.h2 = reinterpret_cast<bf16*>(ws + wl.h2);
  pa.dlogits = reinterpret_cast<bf16*>(ws + wl.dlogits);
  pa.dp1 = reinterpret_cast<bf16*>(ws + wl.dp1);
  pa.dskip = reinterpret_cast<bf16*>(ws + wl.dskip);
  pa.dz = reinterpret_cast<bf16*>(ws + wl.dz);
  pa.B = d.B; pa.T = d.T; pa.D = d.D; pa.S = d.S; pa.P = d.P; pa.Q = d.Q; pa.L = d.L; pa.LD = d.LD;
  pa.use_bias = d.use_bias; pa.rows = d.rows;
  const int maxspq = std::max(d.S, std::max(d.P, d.Q));
  pa.CW = maxspq;
  // K-chunk of the skip GEMM: as many layers as fit next to the widest activation tile
  int lpc = std::max(1, 512 / d.D);
  pa.AW = std::max(maxspq, std::min(d.L, lpc) * d.D);
  return pa;
}

}  // namespace wn

using namespace wn;

extern "C" {

void wn_model_destroy(wn_model* m) {
  if (!m) return;
  if (m->d_layers) cudaFree(m->d_layers);
  if (m->d_kind) cudaFree(m->d_kind);
  delete m;
}

int wn_prof_enable(int32_t on) {
  // 0: off; 1: every category; otherwise bit (c + 1) selects category c (an event pair costs ~1 us of stream time per
  // launch, so bench.py times only the dominant kernel inside its timed region)
  g_prof.mask = on == 0 ? 0u : (on == 1 ? 0xffffffffu : ((uint32_t)on >> 1));
  if (on) {
    g_prof.used = 0;
    g_prof.recs.clear();
  }
  return WN_OK;
}

int wn_prof_collect(double* h_ms, int64_t* h_launches) {
  WN_CUDA_CHECK(cudaDeviceSynchronize());
  for (int i = 0; i < PROF_NCAT; ++i) {
    if (h_ms) h_ms[i] = 0.0;
    if (h_launches) h_launches[i] = 0;
  }
  for (const ProfState::Rec& r : g_prof.recs) {
    float ms = 0.f;
    WN_CUDA_CHECK(cudaEventElapsedTime(&ms, g_prof.pool[r.e0], g_prof.pool[r.e1]));
    if (h_ms) h_ms[r.cat] += ms;
    if (h_launches) h_launches[r.cat] += r.launches;
  }
  g_prof.used = 0;
  g_prof.recs.clear();
  return WN_OK;
}

int wn_debug_trace(void* d_buf, int32_t layer) {
  g_trace_buf = reinterpret_cast<long long*>(d_buf);
  g_trace_layer = layer;
  return WN_OK;
}

int64_t wn_launch_count_reset(void) {
  int64_t n = g_launches;
  g_launches = 0;
  return n;
}

int wn_train_forward(wn_model* m, const float* d_params, void* d_save, const int32_t* d_wav,
                     const int32_t* d_ids, const float* d_mel, int32_t T, void* d_ws, double* d_stats, float* d_logits,
                     void* stream_) {
  if (!m || !d_params || !d_save || !d_wav || !d_ids || !d_ws || !d_stats || T < 2) {
    set_error("wn_train_forward: invalid argument");
    return WN_ERR_INVALID;
  }
  const bool lc = m->a.n_lc_out > 0;
  if (lc && (!d_mel || T % m->lc_hop != 0)) {
    set_error("wn_train_forward: local conditioning needs d_mel and slice_sz %% prod(lc_upsample) == 0");
    return WN_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream_;
  int rc = ensure_tables(m, T);
  if (rc) return rc;
  const WorkspaceLayout& wl = workspace_layout(m, T);
  const Dims d = make_dims(m, T);
  unsigned char* ws = (unsigned char*)d_ws;
  bf16* wbf = reinterpret_cast<bf16*>(ws + wl.wbf);
  const bool gc = d.G > 0;

  const bool umma_post = umma_post_supported(m);
  const bool umma_chain = !umma_post && umma_post_chain_supported(m);
  const bool umma_layer = umma_layer_supported(m);
  const bool umma_wide = !umma_layer && umma_wide_layer_supported(m);
  if (lc && !umma_layer) {
    set_error("wn_train_forward: local conditioning runs on the fused tcgen05 layer kernels only (n_res == n_dil == 32)");
    return WN_ERR_UNSUPPORTED;
  }
  if (!(umma_layer || umma_wide) || !(umma_post || umma_chain) || (umma_wide && !umma_wgrad_x_supported(m, T))) {
    // still the device path, but 10-20x slower than the tcgen05 kernels: say so once instead of silently
    static bool noted = false;
    if (!noted && getenv("WN_DISABLE_UMMA") == nullptr) {
      noted = true;
      fprintf(stderr, "libwavenet_b200: n_res=%d n_dil=%d n_skip=%d n_post=%d slice_sz=%d runs (partly) on the generation-1 "
                      "HMMA kernels; the tcgen05 kernels cover n_res = n_dil = 32 or multiples of 64 (then slice_sz %% 64 == 0), "
                      "n_skip / n_post multiples of 64 up to 512\n", d.R, d.D, d.S, d.P, T);
    }
  }
  {
  ProfScope ps_prep(PROF_PREP, st);
  k_cast_params<<<(unsigned)((m->n_param_elems + 255) / 256), 256, 0, st>>>(d_params, wbf, m->n_param_elems);
  WN_LAUNCH_CHECK();
  if (d.use_bias) {
    k_skip_bias_sum<<<(d.S + 127) / 128, 128, 0, st>>>(d_params, m->d_layers, d.L, d.S,
                                                        reinterpret_cast<float*>(ws + wl.skip_bias));
    WN_LAUNCH_CHECK();
  }
  if (gc) {
    k_gc_table<<<dim3(d.C1, d.L), 2 * d.D, 0, st>>>(d_params, m->off_gc_embed, m->d_layers, d.C1, d.G, d.D,
                                                    reinterpret_cast<float*>(ws + wl.gc_tbl));
    WN_LAUNCH_CHECK();
  }
  if ((umma_post || umma_chain) && (rc = launch_prep_umma(m, d_params, ws, st))) return rc;
  if (umma_layer && (rc = launch_prep_layer_umma(m, d_params, ws, st))) return rc;
  if (umma_wide && (rc = launch_prep_wide_umma(m, d_params, ws, st))) return rc;
  WN_CUDA_CHECK(cudaMemsetAsync(d_stats, 0, sizeof(double) * 3, st));
  WN_CUDA_CHECK(cudaMemsetAsync(ws + wl.tile_ctr, 0, sizeof(int) * 4 * d.L, st));  // tile schedulers (they also re-arm themselves)
  k_save_load<<<dim3(64, d.L), 256, 0, st>>>(reinterpret_cast<const bf16*>(d_save), ws, m->d_layers, d.B, T, d.R);
  WN_LAUNCH_CHECK();
  {
    const int64_t n = d.rows * (d.R / 8);
    k_embed<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_params, m->off_pre, m->off_pre_b, d_wav,
                                                         reinterpret_cast<bf16*>(ws + wl.xfull[0]), d.B, T, d.R,
                                                         m->layers[0].dil, d.Q);
    WN_LAUNCH_CHECK();
  }
  }
  if (lc && (rc = launch_lc_fwd(m, d_params, d_mel, ws, T, st))) return rc;
  const size_t lsm = layer_fwd_smem(d.R, d.D);
  rc = set_smem(k_layer_fwd, lsm);
  if (rc) return rc;
  std::unique_ptr<ProfGroup> pg_fwd(new ProfGroup(PROF_LAYER_FWD, st, umma_layer));
  for (int l = 0; l < d.L; ++l) {
    if (umma_layer) {
      if ((rc = launch_layer_fwd_umma(m, d_params, ws, d_ids, T, l, st))) return rc;
      continue;
    }
    if (umma_wide) {
      if ((rc = launch_layer_fwd_wide_umma(m, d_params, ws, T, l, st))) return rc;
      continue;
    }
    ProfScope ps(PROF_LAYER_FWD, st);
    LayerArgs la;
    la.wbf = wbf; la.params = d_params; la.ld = m->layers[l];
    la.xin = reinterpret_cast<const bf16*>(ws + wl.xfull[l]);
    la.xout = (l + 1 < d.L) ? reinterpret_cast<bf16*>(ws + wl.xfull[l + 1]) : nullptr;
    la.dil_next = (l + 1 < d.L) ? m->layers[l + 1].dil : 0;
    la.z = reinterpret_cast<bf16*>(ws + wl.z);
    la.gc_tbl = gc ? reinterpret_cast<const float*>(ws + wl.gc_tbl) + (size_t)l * d.C1 * 2 * d.D : nullptr;
    la.ids = d_ids;
    la.B = d.B; la.T = T; la.R = d.R; la.D = d.D; la.LD = d.LD; la.l = l; la.C1 = d.C1;
    k_layer_fwd<<<dim3((T + TM - 1) / TM, d.B), NT, lsm, st>>>(la);
    WN_LAUNCH_CHECK();
  }
  pg_fwd.reset();
  {
    ProfScope ps(PROF_PREP, st);
    k_save_store<<<dim3(64, d.L), 256, 0, st>>>(reinterpret_cast<bf16*>(d_save), ws, m->d_layers, d.B, T, d.R);
    WN_LAUNCH_CHECK();
  }
  if (umma_post) return launch_post_fwd_umma(m, d_params, ws, d_wav, d_ids, T, d_stats, d_logits, st);
  if (umma_chain) return launch_post_fwd_chain_umma(m, d_params, ws, d_wav, d_ids, T, d_stats, d_logits, st);
  PostArgs pa = make_post_args(m, d, wl, ws, d_params);
  pa.logits_out = d_logits; pa.wav = d_wav; pa.ids = d_ids; pa.stats = d_stats;
  const size_t psm = post_smem(pa.AW, pa.CW);
  rc = set_smem(k_post_fwd, psm);
  if (rc) return rc;
  {
    ProfScope ps(PROF_POST_FWD, st);
    k_post_fwd<<<(unsigned)((d.rows + TM - 1) / TM), NT, psm, st>>>(pa);
    WN_LAUNCH_CHECK();
  }
  return WN_OK;
}

static int launch_wgrad(const WgradArgs& base, int sm_count, cudaStream_t st) {
  ProfScope ps(PROF_WGRAD, st);
  WgradArgs a = base;
  const int64_t rows = (int64_t)a.B * a.T;
  const int nf = a.N > 64 ? 4 : 1;
  const int n_blk = 64 * nf;
  dim3 grid;
  grid.y = (a.Ka + 31) / 32;
  grid.z = (a.N + n_blk - 1) / n_blk;
  int64_t want = std::max<int64_t>(1, (int64_t)sm_count * 4 / (grid.y * grid.z));
  int64_t chunks = (rows + TM - 1) / TM;
  int64_t split = std::min(chunks, want);
  a.rows_per_cta = (chunks + split - 1) / split * TM;
  grid.x = (unsigned)((rows + a.rows_per_cta - 1) / a.rows_per_cta);
  if (nf == 4)
    k_wgrad<4><<<grid, NT, 0, st>>>(a);
  else
    k_wgrad<1><<<grid, NT, 0, st>>>(a);
  WN_LAUNCH_CHECK();
  return WN_OK;
}

int wn_train_backward(wn_model* m, const float* d_params, const int32_t* d_wav, const int32_t* d_ids,
                      int32_t T, void* d_ws, float* d_grads, void* stream_) {
  return wn_train_backward_phases(m, d_params, d_wav, d_ids, T, d_ws, d_grads, 0, m ? m->L + 2 : 0, stream_);
}

// phases: 0 = zero grads, post-net backward, post-net and SKIP weight gradients;
//         p in [1, L] = layer L-p (gate backward, weight gradients, data gradient);
//         L+1 = PRE gather backward and global-conditioning gradients.
// Gradients of layer l are final once phase L-l has been issued (global-conditioning projections:
// only after phase L+1), which lets data-parallel ranks start all-reducing finished arena ranges
// while earlier layers are still being differentiated.
int wn_train_backward_phases(wn_model* m, const float* d_params, const int32_t* d_wav, const int32_t* d_ids,
                             int32_t T, void* d_ws, float* d_grads, int32_t phase_begin, int32_t phase_end,
                             void* stream_) {
  if (!m || !d_params || !d_wav || !d_ids || !d_ws || !d_grads || T < 2 || phase_begin < 0 ||
      phase_end > m->L + 2 || phase_begin > phase_end) {
    set_error("wn_train_backward: invalid argument");
    return WN_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream_;
  int rc = ensure_tables(m, T);
  if (rc) return rc;
  const WorkspaceLayout& wl = workspace_layout(m, T);
  const Dims d = make_dims(m, T);
  unsigned char* ws = (unsigned char*)d_ws;
  const bf16* wbf = reinterpret_cast<const bf16*>(ws + wl.wbf);
  const bool gc = d.G > 0;
  PostArgs pa = make_post_args(m, d, wl, ws, d_params);
  pa.grads = d_grads;
  if (phase_begin == 0) {
    WN_CUDA_CHECK(cudaMemsetAsync(d_grads, 0, sizeof(float) * m->n_param_elems, st));
    if (gc)
      WN_CUDA_CHECK(cudaMemsetAsync(ws + wl.dgc_tbl, 0, sizeof(float) * (size_t)d.L * d.C1 * 2 * d.D, st));
    if (umma_post_supported(m)) {
      if ((rc = launch_post_bwd_umma(m, ws, T, d_grads, st))) return rc;
    } else if (umma_post_chain_supported(m)) {
      if ((rc = launch_post_bwd_chain_umma(m, ws, T, d_grads, st))) return rc;
    } else {
      const size_t psm = post_smem(pa.AW, pa.CW);
      rc = set_smem(k_post_bwd, psm);
      if (rc) return rc;
      {
        ProfScope ps(PROF_POST_BWD, st);
        k_post_bwd<<<(unsigned)((d.rows + TM - 1) / TM), NT, psm, st>>>(pa);
        WN_LAUNCH_CHECK();
      }
      if (d.use_bias && d.L > 1) {
        k_bcast_skip_bias<<<(d.S + 127) / 128, 128, 0, st>>>(d_grads, m->d_layers, d.L, d.S);
        WN_LAUNCH_CHECK();
      }
    }
  }
  // post-net weight gradients
  WgradArgs wa;
  memset(&wa, 0, sizeof(wa));
  wa.B = d.B; wa.T = T;
  auto flat = [&](const bf16* A, int lda, int a_col0, int Ka, const bf16* Y, int ldy, int y_col0, int N,
                  float* out, int ldo) {
    wa.A = A; wa.a_slot_pitch = (int64_t)T * lda; wa.a_row_off = 0; wa.lda = lda; wa.a_col0 = a_col0; wa.Ka = Ka;
    wa.Y = Y; wa.y_slot_pitch = (int64_t)T * ldy; wa.ldy = ldy; wa.y_col0 = y_col0; wa.N = N;
    wa.out = out; wa.ldo = ldo;
    return launch_wgrad(wa, m->sm_count, st);
  };
  if (phase_begin == 0) {
    if (umma_wgrad_supported(m, d.P, d.Q, d.Q) && umma_wgrad_supported(m, d.S, d.P, d.P) &&
        umma_wgrad_supported(m, d.LD, d.S, d.S)) {
      // POST2 [P][Q] = h2^T dlogits ; POST1 [S][P] = h1^T dp1 ; SKIP_l [D][S] = z_l^T dskip for all layers at once
      if ((rc = launch_wgrad_umma(m, pa.h2, d.P, 0, d.P, pa.dlogits, d.Q, d.Q, d.rows, d_grads + m->off_post2, d.Q, 0,
                                  d_grads, st))) return rc;
      if ((rc = launch_wgrad_umma(m, pa.h1, d.S, 0, d.S, pa.dp1, d.P, d.P, d.rows, d_grads + m->off_post1, d.P, 0,
                                  d_grads, st))) return rc;
      if ((rc = launch_wgrad_umma(m, pa.z, d.LD, 0, d.LD, pa.dskip, d.S, d.S, d.rows, nullptr, d.S, 1, d_grads, st)))
        return rc;
    } else if (umma_post_chain_supported(m)) {
      if ((rc = launch_wgrad_umma_cols(m, pa.h2, d.P, d.P, pa.dlogits, d.Q, d.rows, d_grads + m->off_post2, 0, d_grads, st)))
        return rc;
      if ((rc = launch_wgrad_umma_cols(m, pa.h1, d.S, d.S, pa.dp1, d.P, d.rows, d_grads + m->off_post1, 0, d_grads, st)))
        return rc;
      if ((rc = launch_wgrad_umma_cols(m, pa.z, d.LD, d.LD, pa.dskip, d.S, d.rows, nullptr, 1, d_grads, st))) return rc;
    } else {
      if ((rc = flat(pa.h2, d.P, 0, d.P, pa.dlogits, d.Q, 0, d.Q, d_grads + m->off_post2, d.Q))) return rc;
      if ((rc = flat(pa.h1, d.S, 0, d.S, pa.dp1, d.P, 0, d.P, d_grads + m->off_post1, d.P))) return rc;
      for (int l = 0; l < d.L; ++l)
        if ((rc = flat(pa.z, d.LD, l * d.D, d.D, pa.dskip, d.S, 0, d.S, d_grads + m->layers[l].skip, d.S))) return rc;
    }
  }

  const size_t sa = layer_bwd_a_smem(d.R, d.D), sb = layer_bwd_b_smem(d.R, d.D);
  if ((rc = set_smem(k_layer_bwd_a, sa))) return rc;
  if ((rc = set_smem(k_layer_bwd_b, sb))) return rc;
  bf16* dxbuf[2] = {reinterpret_cast<bf16*>(ws + wl.dx[0]), reinterpret_cast<bf16*>(ws + wl.dx[1])};
  bf16* dv = reinterpret_cast<bf16*>(ws + wl.dv);
  const bool fused = umma_bwd_fused_supported(m);  // data gradient in split form dx[t] = Y[t] + P0[t + dil]
  std::unique_ptr<ProfGroup> pg_bwd(new ProfGroup(PROF_LAYER_BWD_A, st, fused));
  for (int l = d.L - 1; l >= 0; --l) {
    const int phase = d.L - l;
    if (phase < phase_begin || phase >= phase_end) continue;
    const bf16* dx_next = (l == d.L - 1) ? nullptr : dxbuf[(l + 1) & 1];
    LayerBwdArgs la;
    memset(&la, 0, sizeof(la));
    la.wbf = wbf; la.params = d_params; la.ld = m->layers[l];
    la.xin = reinterpret_cast<const bf16*>(ws + wl.xfull[l]);
    la.dx_next = dx_next;
    la.dx_out = dxbuf[l & 1];
    la.dz = pa.dz; la.dv = dv;
    la.gc_tbl = gc ? reinterpret_cast<const float*>(ws + wl.gc_tbl) + (size_t)l * d.C1 * 2 * d.D : nullptr;
    la.dgc_tbl = gc ? reinterpret_cast<float*>(ws + wl.dgc_tbl) + (size_t)l * d.C1 * 2 * d.D : nullptr;
    la.ids = d_ids; la.grads = d_grads;
    la.B = d.B; la.T = T; la.R = d.R; la.D = d.D; la.LD = d.LD; la.l = l; la.C1 = d.C1;
    const dim3 grid((T + TM - 1) / TM, d.B);
    if (fused) {
      // gate backward, conv / residual weight + bias gradients and the data gradient in one persistent tcgen05 kernel
      if ((rc = launch_layer_bwd_fused_umma(m, d_params, ws, d_ids, T, l, d_grads, st))) return rc;
      continue;
    }
    const bool wide = umma_wide_layer_supported(m) && umma_wgrad_x_supported(m, T);
    if (wide) {
      if ((rc = launch_layer_bwd_wide_umma(m, d_params, ws, T, l, d_grads, st))) return rc;
    } else {
      ProfScope ps(PROF_LAYER_BWD_A, st);
      k_layer_bwd_a<<<grid, NT, sa, st>>>(la);
      WN_LAUNCH_CHECK();
    }
    // weight gradients of this layer
    const bool wg_umma = umma_wgrad_x_supported(m, T);  // tcgen05 split-K kernel (MN-major operands)
    if (dx_next != nullptr) {  // RESIDUAL_l [D][R] = z_l^T dx_{l+1}  (the last layer's output is unused)
      if (wg_umma) {
        if ((rc = launch_wgrad_umma(m, pa.z, d.LD, l * d.D, d.D, dx_next, d.R, d.R, d.rows, d_grads + m->layers[l].res,
                                    d.R, 0, d_grads, st))) return rc;
      } else if ((rc = flat(pa.z, d.LD, l * d.D, d.D, dx_next, d.R, 0, d.R, d_grads + m->layers[l].res, d.R))) {
        return rc;
      }
    }
    if (wg_umma && 2 * d.D <= 256) {
      // SIGNAL and GATE gradients of both taps in one launch (N = 2D with a split output, tap = blockIdx.x / m_tiles)
      if ((rc = launch_wgrad_umma_x(m, la.xin, la.ld.dil, T, -1, dv, 2 * d.D, 2 * d.D, d_grads + la.ld.sig,
                                    d_grads + la.ld.gate, d.D, d.D, (int64_t)d.R * d.D, st)))
        return rc;
    } else
    for (int tap = 0; tap < 2; ++tap) {
      for (int sg = 0; sg < 2; ++sg) {
        if (wg_umma) {
          if ((rc = launch_wgrad_umma_x(m, la.xin, la.ld.dil, T, tap, dv + sg * d.D, 2 * d.D, d.D,
                                        d_grads + (sg ? la.ld.gate : la.ld.sig) + (int64_t)tap * d.R * d.D, nullptr, 0,
                                        d.D, 0, st)))
            return rc;
          continue;
        }
        wa.A = la.xin; wa.a_slot_pitch = (int64_t)(la.ld.dil + T) * d.R; wa.a_row_off = tap ? la.ld.dil : 0;
        wa.lda = d.R; wa.a_col0 = 0; wa.Ka = d.R;
        wa.Y = dv; wa.y_slot_pitch = (int64_t)T * 2 * d.D; wa.ldy = 2 * d.D; wa.y_col0 = sg * d.D; wa.N = d.D;
        wa.out = d_grads + (sg ? la.ld.gate : la.ld.sig) + (int64_t)tap * d.R * d.D; wa.ldo = d.D;
        if ((rc = launch_wgrad(wa, m->sm_count, st))) return rc;
      }
    }
    if (!wide) {
      ProfScope ps(PROF_LAYER_BWD_B, st);
      k_layer_bwd_b<<<grid, NT, sb, st>>>(la);
      WN_LAUNCH_CHECK();
    }
  }
  pg_bwd.reset();
  if (phase_end < d.L + 2) return WN_OK;
  ProfScope ps_tail(PROF_EMBED_GC_BWD, st);
  {
    const bf16* dx_next = dxbuf[0];  // gradient wrt the layer-0 input
    const int ncopy = 1;  // (private table copies did not pay: the kernel is load-latency bound, not atomic bound)
    const size_t esm = (size_t)ncopy * (d.Q + 1) * d.R * sizeof(float);
    if ((rc = set_smem(k_embed_bwd, esm))) return rc;
    // two 1024-thread CTAs per SM when the table is small enough: twice the loads in flight
    const int per_sm = esm <= 96 * 1024 ? 2 : 1;
    const int nblk = (int)std::min<int64_t>(std::min(per_sm * m->sm_count, WN_EMBED_PARTS), (d.rows + 1023) / 1024);
    const int64_t rpc = (d.rows + nblk - 1) / nblk;
    const bf16* p0 = nullptr;  // (the fused backward hands over the merged dx_0 as every other path does)
    float* part = reinterpret_cast<float*>(ws + wl.embed_part);
    k_embed_bwd<<<nblk, 1024, esm, st>>>(dx_next, p0, m->layers[0].dil, T, d_wav, part, d.rows, d.R, d.Q, rpc, ncopy);
    WN_LAUNCH_CHECK();
    k_embed_bwd_reduce<<<((d.Q + 1) * d.R + 255) / 256, 256, 0, st>>>(part, nblk, d_grads, m->off_pre, m->off_pre_b, d.R, d.Q);
    WN_LAUNCH_CHECK();
  }
  if (m->a.n_lc_out > 0 && (rc = launch_lc_bwd(m, ws, T, d_grads, st))) return rc;
  if (gc) {
    const float* dtbl = reinterpret_cast<const float*>(ws + wl.dgc_tbl);
    k_gc_bwd_embed<<<d.C1, 64, 0, st>>>(d_params, dtbl, m->d_layers, d_grads, m->off_gc_embed, d.L, d.C1, d.G, d.D);
    WN_LAUNCH_CHECK();
    k_gc_bwd_proj<<<dim3(d.G, d.L), 2 * d.D, 0, st>>>(d_params, dtbl, m->d_layers, d_grads, m->off_gc_embed,
                                                       d.C1, d.G, d.D);
    WN_LAUNCH_CHECK();
  }
  return WN_OK;
}

int wn_adam_step(wn_model* m, float* d_params, const float* d_grads, float* d_m, float* d_v,
                 const double* d_n_valid, int32_t step, float lr, float l2_factor, float beta1, float beta2,
                 float eps, void* stream_) {
  if (!m || !d_params || !d_grads || !d_m || !d_v || !d_n_valid || step < 1) {
    set_error("wn_adam_step: invalid argument");
    return WN_ERR_INVALID;
  }
  int rc = ensure_layer_table(m);
  if (rc) return rc;
  if ((rc = ensure_kind(m))) return rc;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step));
  const int64_t n = m->n_param_elems;
  ProfScope ps(PROF_ADAM, (cudaStream_t)stream_);
  k_adam<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
      d_params, d_grads, d_m, d_v, m->d_kind, d_n_valid, n, (float)lr_t, l2_factor, beta1, beta2, eps);
  WN_LAUNCH_CHECK();
  return WN_OK;
}

int wn_l2_loss(wn_model* m, const float* d_params, double* d_stats, void* stream_) {
  if (!m || !d_params || !d_stats) {
    set_error("wn_l2_loss: invalid argument");
    return WN_ERR_INVALID;
  }
  int rc = ensure_layer_table(m);
  if (rc) return rc;
  if ((rc = ensure_kind(m))) return rc;
  cudaStream_t st = (cudaStream_t)stream_;
  WN_CUDA_CHECK(cudaMemsetAsync(d_stats + WN_STAT_L2, 0, sizeof(double), st));
  k_l2<<<std::max(1, m->sm_count), NT, 0, st>>>(d_params, m->d_kind, m->n_param_elems, d_stats + WN_STAT_L2);
  WN_LAUNCH_CHECK();
  return WN_OK;
}

int wn_codes_u8_to_i32(const uint8_t* d_u8, int32_t* d_i32, int64_t n, void* stream_) {
  if (!d_u8 || !d_i32 || n < 0) {
    set_error("wn_codes_u8_to_i32: invalid argument");
    return WN_ERR_INVALID;
  }
  if (n == 0) return WN_OK;
  const int64_t n16 = (n + 15) / 16;
  k_widen_u8<<<(unsigned)std::min<int64_t>((n16 + 255) / 256, 148 * 16), 256, 0, (cudaStream_t)stream_>>>(d_u8, d_i32, n);
  WN_LAUNCH_CHECK();
  return WN_OK;
}

int wn_debug_read(wn_model* m, const void* d_ws, int32_t T, int32_t what, int32_t layer, float* d_out,
                  void* stream_) {
  if (!m || !d_ws || !d_out) {
    set_error("wn_debug_read: invalid argument");
    return WN_ERR_INVALID;
  }
  const WorkspaceLayout& wl = workspace_layout(m, T);
  const Dims d = make_dims(m, T);
  const unsigned char* ws = (const unsigned char*)d_ws;
  const bf16* src = nullptr;
  const bf16* add = nullptr;
  int ncols = 0, ld = 0, col0 = 0, row_off = 0, add_shift = 0;
  int64_t slot_rows = T, slot_pitch_rows = T;
  switch (what) {
    case 0:
      if (layer < 0 || layer >= d.L) { set_error("wn_debug_read: bad layer"); return WN_ERR_INVALID; }
      src = reinterpret_cast<const bf16*>(ws + wl.xfull[layer]);
      ncols = d.R; ld = d.R; row_off = m->layers[layer].dil; slot_pitch_rows = m->layers[layer].dil + T;
      break;
    case 1:
      if (layer < 0 || layer >= d.L) { set_error("wn_debug_read: bad layer"); return WN_ERR_INVALID; }
      src = reinterpret_cast<const bf16*>(ws + wl.z); ncols = d.D; ld = d.LD; col0 = layer * d.D;
      break;
    case 2: src = reinterpret_cast<const bf16*>(ws + wl.h1); ncols = d.S; ld = d.S; break;
    case 3: src = reinterpret_cast<const bf16*>(ws + wl.h2); ncols = d.P; ld = d.P; break;
    case 4: src = reinterpret_cast<const bf16*>(ws + wl.dlogits); ncols = d.Q; ld = d.Q; break;
    case 5:  // gradient wrt the layer-0 input
      src = reinterpret_cast<const bf16*>(ws + wl.dx[0]); ncols = d.R; ld = d.R;
      break;
    case 6:  // dz plane of `layer` ([L][B*T][D]): the skip-path gradient after phase 0 (the wide-layer backward adds the
             // residual part in place while it differentiates that layer)
      if (layer < 0 || layer >= d.L) { set_error("wn_debug_read: bad layer"); return WN_ERR_INVALID; }
      src = reinterpret_cast<const bf16*>(ws + wl.dz) + (size_t)layer * d.rows * d.D; ncols = d.D; ld = d.D;
      break;
    case 7:  // data-gradient buffer of parity `layer` & 1: dx_l right after layer l's backward
      src = reinterpret_cast<const bf16*>(ws + wl.dx[layer & 1]); ncols = d.R; ld = d.R;
      break;
    case 9:  // local-conditioning plane of `layer` [., ., 2 n_dil]: the projections lc_up . [LC_SIGNAL_l | LC_GATE_l]
      if (layer < 0 || layer >= d.L || m->a.n_lc_out == 0) { set_error("wn_debug_read: bad layer / no local conditioning"); return WN_ERR_INVALID; }
      src = reinterpret_cast<const bf16*>(ws + wl.cond) + (size_t)layer * d.rows * 2 * d.D; ncols = 2 * d.D; ld = 2 * d.D;
      break;
    case 11:  // gradient wrt that plane (dv_l), after layer `layer`'s backward
      if (layer < 0 || layer >= d.L || m->a.n_lc_out == 0) { set_error("wn_debug_read: bad layer / no local conditioning"); return WN_ERR_INVALID; }
      src = reinterpret_cast<const bf16*>(ws + wl.dcond) + (size_t)layer * d.rows * 2 * d.D; ncols = 2 * d.D; ld = 2 * d.D;
      break;
    case 10:  // upsampled local conditioning [., ., 128] (channels >= n_lc_out are zero)
      if (m->a.n_lc_out == 0) { set_error("wn_debug_read: no local conditioning"); return WN_ERR_INVALID; }
      src = reinterpret_cast<const bf16*>(ws + wl.lc_x[m->a.n_lc_layers]); ncols = 128; ld = 128;
      break;
    default: set_error("wn_debug_read: unknown tap %d", what); return WN_ERR_INVALID;
  }
  const int64_t n = d.rows * ncols;
  k_debug_read<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(src, d_out, d.rows, ncols, slot_rows,
                                                                             slot_pitch_rows, row_off, ld, col0, add, add_shift);
  WN_LAUNCH_CHECK();
  return WN_OK;
}

}  // extern "C"
