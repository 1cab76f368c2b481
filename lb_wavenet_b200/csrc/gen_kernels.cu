// Incremental generator + primitive ops (mu-law, sampler).
//
// Reference semantics: imodel.py:61-272 (one timestep of WaveNetGen._loop_body for batch_sz
// streams), ops.py:4-39 (mu-law), imodel.py:179 (sampling; replaced by a seeded
// counter-based sampler, see oracle/wavenet_oracle.py sample_from_logits for the canonical
// evaluation order reproduced here bit for bit).
//
// Generation 1 of the generator kernel: one persistent CTA per group of GS streams walks the
// whole layer stack for every timestep without leaving the kernel (no per-timestep launch,
// no grid-wide sync: streams are independent).  Per-layer ring buffers (length dil, index
// t mod dil) replace the reference's chunk-shifted lookback buffers (imodel.py:88-97,199-201).
#include "common.cuh"
#include "sampler.cuh"

#include <cstring>
#include <vector>

namespace wn {

// generation 2 (gen_mma.cu)
bool gen2_supported(const wn_model* m);
int64_t gen2_blob_bytes(const wn_model* m);
int gen2_prepare(wn_model* m, const float* d_params, unsigned char* blob, cudaStream_t st);
int gen2_run(wn_model* m, const unsigned char* blob, const int64_t* ring_off, bf16* rings, int32_t* codes,
             int n_streams, int64_t t0, int n_steps, uint64_t seed, const int32_t* teacher, int n_teacher, int32_t* out,
             float* logits, const float* gcproj, cudaStream_t st);

__constant__ uint32_t c_mu_thr[255];
__constant__ uint32_t c_mu_dec[256];
static bool g_tables_uploaded = false;

static int upload_tables() {
  if (g_tables_uploaded) return WN_OK;
  WN_CUDA_CHECK(cudaMemcpyToSymbol(c_mu_thr, kMuEncodeThrBits, sizeof(kMuEncodeThrBits)));
  WN_CUDA_CHECK(cudaMemcpyToSymbol(c_mu_dec, kMuDecodeBits, sizeof(kMuDecodeBits)));
  g_tables_uploaded = true;
  return WN_OK;
}

// ---- mu-law ---------------------------------------------------------------------------
// encode(x) = #{q : thr[q] <= x}: the float32 numpy encoder (ops.py:23-28) is monotone, so
// counting thresholds reproduces it on every float32 input in [-1, 1].
__global__ void k_mu_encode(const float* __restrict__ x, int32_t* __restrict__ q, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = x[i];
  int lo = 0, hi = 255;  // number of thresholds <= v, by binary search over the sorted table
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (__uint_as_float(c_mu_thr[mid]) <= v)
      lo = mid + 1;
    else
      hi = mid;
  }
  q[i] = lo;
}

__global__ void k_mu_decode(const int32_t* __restrict__ q, float* __restrict__ x, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int c = min(max(q[i], 0), 255);
  x[i] = __uint_as_float(c_mu_dec[c]);
}

__global__ void k_sample_logits(const float* __restrict__ logits, int n_rows, uint64_t seed, uint64_t step,
                                int32_t* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const float u = sampler_uniform(seed, step, (uint32_t)row);
  const int s = warp_sample(logits + (size_t)row * 256, u);
  if ((threadIdx.x & 31) == 0) out[row] = s;
}

// ---- generator --------------------------------------------------------------------------
constexpr int GS = 4;     // streams per CTA
constexpr int GNT = 256;  // threads per CTA

struct GenLayout {
  int64_t ring_off;  // int64[L]: element offset of every layer's ring
  int64_t codes;     // int32[n_streams] pending input code (-1 == all-zero vector)
  int64_t wbf;       // bf16 mirror of the parameter arena
  int64_t pf32;      // fp32 copy of the arena (PRE table, biases)
  int64_t gcproj;    // fp32 [n_streams][L][2D]
  int64_t blob2;     // generation-2 fragment-ready weight blob (gen_mma.cu)
  int64_t rings;     // bf16, layer l: [n_streams][dil][R]
  int64_t ring_elems;
  int64_t total;
};

static GenLayout gen_layout(const wn_model* m, int n_streams) {
  GenLayout g;
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    int64_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  g.ring_off = take(sizeof(int64_t) * m->L);
  g.codes = take(sizeof(int32_t) * n_streams);
  g.wbf = take(m->n_param_elems * 2);
  g.pf32 = take(m->n_param_elems * 4);
  g.gcproj = take(m->a.n_gc_embed > 0 ? (int64_t)n_streams * m->L * 2 * m->a.n_dil * 4 : 0);
  g.blob2 = take(gen2_supported(m) ? gen2_blob_bytes(m) : 0);
  int64_t e = 0;
  for (int l = 0; l < m->L; ++l) e += (int64_t)n_streams * m->layers[l].dil * m->a.n_res;
  g.ring_elems = e;
  g.rings = take(e * 2);
  g.total = off;
  return g;
}

struct GenArgs {
  const bf16* wbf;
  const float* params;  // fp32 arena (biases, PRE table)
  const LayerDesc* layers;
  const int64_t* ring_off;
  bf16* rings;
  const float* gcproj;
  int32_t* codes;
  const int32_t* teacher;
  int32_t* out;
  float* logits_out;
  int64_t off_pre, off_pre_b, off_post1, off_post1_b, off_post2, off_post2_b;
  int64_t t0;
  uint64_t seed;
  int n_streams, n_steps, n_teacher;
  int R, D, S, P, Q, L;
};

__global__ void __launch_bounds__(GNT) k_gen(GenArgs a) {
  extern __shared__ float gsm[];
  const int R = a.R, D = a.D, S = a.S, P = a.P, Q = a.Q;
  float* x_s = gsm;                 // [GS][R]
  float* old_s = x_s + GS * R;      // [GS][R]
  float* v_s = old_s + GS * R;      // [GS][2D]
  float* z_s = v_s + GS * 2 * D;    // [GS][D]
  float* skip_s = z_s + GS * D;     // [GS][S]
  float* h_s = skip_s + GS * S;     // [GS][S]
  float* h2_s = h_s + GS * S;       // [GS][P]
  float* lg_s = h2_s + GS * P;      // [GS][Q]
  __shared__ int code_s[GS];
  const int s0 = blockIdx.x * GS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid < GS) code_s[tid] = (s0 + tid < a.n_streams) ? a.codes[s0 + tid] : -1;
  __syncthreads();
  const float* p = a.params;
  for (int step = 0; step < a.n_steps; ++step) {
    const int64_t t = a.t0 + step;
    // PRE: one-hot @ PRE == row gather; all-zero input vector -> bias only (imodel.py:66-74)
    for (int idx = tid; idx < GS * R; idx += GNT) {
      const int s = idx / R, r = idx % R;
      const int c = code_s[s];
      float v = (c >= 0 && c < Q) ? p[a.off_pre + (int64_t)c * R + r] : 0.f;
      if (a.off_pre_b >= 0) v += p[a.off_pre_b + r];
      x_s[idx] = bf2f(f2bf(v));
    }
    for (int idx = tid; idx < GS * S; idx += GNT) skip_s[idx] = 0.f;
    __syncthreads();
    for (int l = 0; l < a.L; ++l) {
      const LayerDesc ld = a.layers[l];
      const int slot = (int)(t % ld.dil);
      bf16* ring = a.rings + a.ring_off[l];
      for (int idx = tid; idx < GS * R; idx += GNT) {
        const int s = idx / R, r = idx % R;
        if (s0 + s < a.n_streams) {
          bf16* cell = ring + ((int64_t)(s0 + s) * ld.dil + slot) * R + r;
          old_s[idx] = bf2f(*cell);   // x[t-dil]            (imodel.py:107)
          *cell = f2bf(x_s[idx]);     // ring <- x[t]        (imodel.py:97)
        } else {
          old_s[idx] = 0.f;
        }
      }
      __syncthreads();
      for (int idx = tid; idx < GS * 2 * D; idx += GNT) {
        const int s = idx / (2 * D), n = idx % (2 * D);
        const bool gate = n >= D;
        const int d = gate ? n - D : n;
        const bf16* W = a.wbf + (gate ? ld.gate : ld.sig) + d;  // [2][R][D]
        float acc = 0.f;
        const float* xo = old_s + s * R;
        const float* xc = x_s + s * R;
        for (int k = 0; k < R; ++k) acc += xo[k] * bf2f(W[(int64_t)k * D]);
        for (int k = 0; k < R; ++k) acc += xc[k] * bf2f(W[(int64_t)(R + k) * D]);
        const int64_t bo = gate ? ld.gate_b : ld.sig_b;
        if (bo >= 0) acc += p[bo + d];
        if (a.gcproj != nullptr && s0 + s < a.n_streams)
          acc += a.gcproj[((int64_t)(s0 + s) * a.L + l) * 2 * D + n];
        v_s[idx] = acc;
      }
      __syncthreads();
      for (int idx = tid; idx < GS * D; idx += GNT) {
        const int s = idx / D, d = idx % D;
        z_s[idx] = bf2f(f2bf(tanh_fast(v_s[s * 2 * D + d]) * sigmoid_fast(v_s[s * 2 * D + D + d])));
      }
      __syncthreads();
      for (int idx = tid; idx < GS * (R + S); idx += GNT) {
        const int s = idx / (R + S), j = idx % (R + S);
        const float* zz = z_s + s * D;
        if (j < R) {
          const bf16* W = a.wbf + ld.res + j;  // [D][R]
          float acc = 0.f;
          for (int k = 0; k < D; ++k) acc += zz[k] * bf2f(W[(int64_t)k * R]);
          if (ld.res_b >= 0) acc += p[ld.res_b + j];
          x_s[s * R + j] = bf2f(f2bf(x_s[s * R + j] + acc));  // imodel.py:245
        } else {
          const int c = j - R;
          const bf16* W = a.wbf + ld.skip + c;  // [D][S]
          float acc = 0.f;
          for (int k = 0; k < D; ++k) acc += zz[k] * bf2f(W[(int64_t)k * S]);
          if (ld.skip_b >= 0) acc += p[ld.skip_b + c];
          skip_s[s * S + c] += acc;  // imodel.py:247
        }
      }
      __syncthreads();
    }
    // post-net (imodel.py:140-164)
    for (int idx = tid; idx < GS * S; idx += GNT) h_s[idx] = bf2f(f2bf(fmaxf(skip_s[idx], 0.f)));
    __syncthreads();
    for (int idx = tid; idx < GS * P; idx += GNT) {
      const int s = idx / P, c = idx % P;
      const bf16* W = a.wbf + a.off_post1 + c;
      const float* hh = h_s + s * S;
      float acc = 0.f;
      for (int k = 0; k < S; ++k) acc += hh[k] * bf2f(W[(int64_t)k * P]);
      if (a.off_post1_b >= 0) acc += p[a.off_post1_b + c];
      h2_s[s * P + c] = bf2f(f2bf(fmaxf(acc, 0.f)));
    }
    __syncthreads();
    for (int idx = tid; idx < GS * Q; idx += GNT) {
      const int s = idx / Q, c = idx % Q;
      const bf16* W = a.wbf + a.off_post2 + c;
      const float* hh = h2_s + s * P;
      float acc = 0.f;
      for (int k = 0; k < P; ++k) acc += hh[k] * bf2f(W[(int64_t)k * Q]);
      if (a.off_post2_b >= 0) acc += p[a.off_post2_b + c];
      lg_s[idx] = acc;
      if (a.logits_out != nullptr && s0 + s < a.n_streams)
        a.logits_out[((int64_t)(s0 + s) * a.n_steps + step) * Q + c] = acc;
    }
    __syncthreads();
    if (warp < GS && s0 + warp < a.n_streams) {
      const int s = warp;
      const float u = sampler_uniform(a.seed, (uint64_t)t, (uint32_t)(s0 + s));
      const int samp = warp_sample(lg_s + s * Q, u);  // imodel.py:179
      if (lane == 0) {
        a.out[(int64_t)(s0 + s) * a.n_steps + step] = samp;
        code_s[s] = (t < a.n_teacher) ? a.teacher[t] : samp;  // imodel.py:260-267
      }
    }
    __syncthreads();
  }
  if (tid < GS && s0 + tid < a.n_streams) a.codes[s0 + tid] = code_s[tid];
}

__global__ void k_gen_gcproj(const float* __restrict__ p, int64_t off_embed, const LayerDesc* __restrict__ layers,
                             const int32_t* __restrict__ gc_ids, int C1, int G, int D, int L,
                             float* __restrict__ out) {
  const int s = blockIdx.x, l = blockIdx.y, n = threadIdx.x;
  if (n >= 2 * D) return;
  int c = gc_ids ? gc_ids[s] : 0;
  c = min(max(c, 0), C1 - 1);
  const float* e = p + off_embed + (int64_t)c * G;
  const float* w = n < D ? p + layers[l].gc_sig + n : p + layers[l].gc_gate + (n - D);
  float acc = 0.f;
  for (int g = 0; g < G; ++g) acc += e[g] * w[(int64_t)g * D];
  out[((int64_t)s * L + l) * 2 * D + n] = acc;
}

__global__ void k_fill_i32(int32_t* p, int32_t v, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void k_cast_params_gen(const float* __restrict__ p, bf16* __restrict__ w, float* __restrict__ c,
                                  int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float v = p[i];
    w[i] = f2bf(v);
    c[i] = v;
  }
}

int ensure_layer_table(wn_model* m);  // train_kernels.cu


}  // namespace wn

using namespace wn;

extern "C" {

int wn_mu_encode(const float* d_x, int32_t* d_q, int64_t n, void* stream) {
  int rc = upload_tables();
  if (rc) return rc;
  if (n <= 0) return WN_OK;
  k_mu_encode<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_x, d_q, n);
  WN_LAUNCH_CHECK();
  return WN_OK;
}

int wn_mu_decode(const int32_t* d_q, float* d_x, int64_t n, void* stream) {
  int rc = upload_tables();
  if (rc) return rc;
  if (n <= 0) return WN_OK;
  k_mu_decode<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(d_q, d_x, n);
  WN_LAUNCH_CHECK();
  return WN_OK;
}

int wn_sample_logits(const float* d_logits, int32_t n_rows, uint64_t seed, int64_t step, int32_t* d_out,
                     void* stream) {
  if (n_rows <= 0) return WN_OK;
  k_sample_logits<<<(n_rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(d_logits, n_rows, seed, (uint64_t)step, d_out);
  WN_LAUNCH_CHECK();
  return WN_OK;
}

int64_t wn_gen_workspace_bytes(const wn_model* m, int32_t n_streams) {
  if (!m || n_streams < 1) {
    set_error("wn_gen_workspace_bytes: invalid argument");
    return WN_ERR_INVALID;
  }
  return gen_layout(m, n_streams).total;
}

int wn_gen_reset(wn_model* m, void* d_gws, int32_t n_streams, void* stream_) {
  if (!m || !d_gws || n_streams < 1) {
    set_error("wn_gen_reset: invalid argument");
    return WN_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream_;
  const GenLayout g = gen_layout(m, n_streams);
  unsigned char* ws = (unsigned char*)d_gws;
  WN_CUDA_CHECK(cudaMemsetAsync(ws + g.rings, 0, g.ring_elems * 2, st));  // imodel.py:88-95
  k_fill_i32<<<(n_streams + 255) / 256, 256, 0, st>>>(reinterpret_cast<int32_t*>(ws + g.codes), -1, n_streams);
  WN_LAUNCH_CHECK();
  std::vector<int64_t> ro(m->L);
  int64_t e = 0;
  for (int l = 0; l < m->L; ++l) {
    ro[l] = e;
    e += (int64_t)n_streams * m->layers[l].dil * m->a.n_res;
  }
  WN_CUDA_CHECK(cudaMemcpyAsync(ws + g.ring_off, ro.data(), sizeof(int64_t) * m->L, cudaMemcpyHostToDevice, st));
  WN_CUDA_CHECK(cudaStreamSynchronize(st));  // ro is a host temporary
  return WN_OK;
}

int wn_gen_load_params(wn_model* m, const float* d_params, const int32_t* d_gc_ids, void* d_gws,
                       int32_t n_streams, void* stream_) {
  if (!m || !d_params || !d_gws || n_streams < 1) {
    set_error("wn_gen_load_params: invalid argument");
    return WN_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream_;
  int rc = ensure_layer_table(m);
  if (rc) return rc;
  const GenLayout g = gen_layout(m, n_streams);
  unsigned char* ws = (unsigned char*)d_gws;
  const int64_t n = m->n_param_elems;
  k_cast_params_gen<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_params, reinterpret_cast<bf16*>(ws + g.wbf),
                                                                       reinterpret_cast<float*>(ws + g.pf32), n);
  WN_LAUNCH_CHECK();
  if (gen2_supported(m) && (rc = gen2_prepare(m, d_params, ws + g.blob2, st))) return rc;
  if (m->a.n_gc_embed > 0) {
    k_gen_gcproj<<<dim3(n_streams, m->L), 2 * m->a.n_dil, 0, st>>>(
        d_params, m->off_gc_embed, m->d_layers, d_gc_ids, m->a.n_gc_category + 1, m->a.n_gc_embed, m->a.n_dil,
        m->L, reinterpret_cast<float*>(ws + g.gcproj));
    WN_LAUNCH_CHECK();
  }
  return WN_OK;
}

int wn_gen_run(wn_model* m, void* d_gws, int32_t n_streams, int64_t t0, int32_t n_steps, uint64_t seed,
               const int32_t* d_teacher, int32_t n_teacher, int32_t* d_out, float* d_logits, void* stream_) {
  if (!m || !d_gws || !d_out || n_streams < 1 || n_steps < 1 || t0 < 0) {
    set_error("wn_gen_run: invalid argument");
    return WN_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream_;
  int rc = ensure_layer_table(m);
  if (rc) return rc;
  const GenLayout g = gen_layout(m, n_streams);
  unsigned char* ws = (unsigned char*)d_gws;
  if (gen2_supported(m))
    return gen2_run(m, ws + g.blob2, reinterpret_cast<const int64_t*>(ws + g.ring_off),
                    reinterpret_cast<bf16*>(ws + g.rings), reinterpret_cast<int32_t*>(ws + g.codes), n_streams, t0,
                    n_steps, seed, d_teacher, d_teacher ? n_teacher : 0, d_out, d_logits,
                    m->a.n_gc_embed > 0 ? reinterpret_cast<const float*>(ws + g.gcproj) : nullptr, st);
  GenArgs a;
  memset(&a, 0, sizeof(a));
  a.wbf = reinterpret_cast<const bf16*>(ws + g.wbf);
  a.layers = m->d_layers;
  a.ring_off = reinterpret_cast<const int64_t*>(ws + g.ring_off);
  a.rings = reinterpret_cast<bf16*>(ws + g.rings);
  a.gcproj = m->a.n_gc_embed > 0 ? reinterpret_cast<const float*>(ws + g.gcproj) : nullptr;
  a.codes = reinterpret_cast<int32_t*>(ws + g.codes);
  a.teacher = d_teacher;
  a.n_teacher = d_teacher ? n_teacher : 0;
  a.out = d_out;
  a.logits_out = d_logits;
  a.off_pre = m->off_pre; a.off_pre_b = m->off_pre_b;
  a.off_post1 = m->off_post1; a.off_post1_b = m->off_post1_b;
  a.off_post2 = m->off_post2; a.off_post2_b = m->off_post2_b;
  a.t0 = t0; a.seed = seed; a.n_streams = n_streams; a.n_steps = n_steps;
  a.R = m->a.n_res; a.D = m->a.n_dil; a.S = m->a.n_skip; a.P = m->a.n_post; a.Q = m->a.n_quant; a.L = m->L;
  // fp32 values (PRE table, biases) come from the copy captured by wn_gen_load_params
  a.params = reinterpret_cast<const float*>(ws + g.pf32);
  const size_t smem = sizeof(float) * (size_t)GS * (2 * a.R + 2 * a.D + a.D + 2 * a.S + a.P + a.Q);
  if (smem > 48 * 1024)
    WN_CUDA_CHECK(cudaFuncSetAttribute(k_gen, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_gen<<<(n_streams + GS - 1) / GS, GNT, smem, st>>>(a);
  WN_LAUNCH_CHECK();
  return WN_OK;
}

}  // extern "C"
