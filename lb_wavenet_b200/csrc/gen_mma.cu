// Incremental generator, generation 2 (reference imodel.py:214-272 for batch_sz independent streams).
//
// One persistent CTA owns 16 streams (the M of mma.sync.m16n8k16) and walks the whole stack for every
// timestep without leaving the kernel.  The per-sample dependency chain (30 layers -> post-net -> sampler)
// is latency bound, so everything that does not depend on the sample is taken off it:
//   * all weights stream through shared memory in FRAGMENT-READY order (every lane fetches its B fragment
//     with one conflict-free 8-byte load) via cp.async.bulk into a 5-slot ring fed by a producer warp;
//     the sequence is the same every timestep, so the producer runs arbitrarily far ahead;
//   * x[t-dil] of every layer is prefetched from the HBM ring buffers ONE STEP AHEAD with cp.async
//     (addresses depend on t only); dil == 1 layers keep their previous input in shared memory;
//   * skip accumulators stay in registers across the 30 layers (each warp owns 32 skip channels);
//   * the input embedding bf16(PRE[code] + bias) is a 16 KB shared-memory table.
// 8 compute warps + 1 producer warp.  Two named barriers per layer.
// Shapes: R = D = 32, Q = 256, S and P in {256, 512}, with or without global conditioning -- i.e. the classic 3x10
// stack that is benchmarked AND every R = D = 32 architecture the reference ships (par/arch1.json, arch3.json:
// S = P = 512, arch1 with a 17-wide voice embedding).  Global conditioning (imodel.py:53-56,113-118) is a per-stream,
// per-layer constant: the projections gc_embed[id] . GC_SIGNAL|GATE_l are computed once per launch sequence
// (wn_gen_load_params) and read as a per-row bias in the gate, prefetched one layer ahead.  Other shapes: k_gen.
#include <algorithm>
#include <cstring>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "sampler.cuh"
#include "umma.cuh"

namespace wn {

using namespace umma;

namespace g2 {
constexpr int GS = 16;            // streams per CTA
constexpr int R = 32, D = 32, Q = 256;   // S, P: template parameters of k_gen2 (256 or 512)
constexpr int NCW = 8;            // compute warps
constexpr int THREADS = (NCW + 1) * 32;
// weight ring slot: one layer item (conv | residual | biases | SKIP_l) or one post-net K slice; S = 256: 5 slots of
// 27 KB, S or P = 512: 3 slots of 43 KB (the 512-wide activation tiles need the rest of the 227 KB)
constexpr int slot_bytes(int S, int P) { return ((10752 + (S > P ? S : P) * 64 + 1023) / 1024) * 1024; }
constexpr int n_slots(int S, int P) { return (S > 256 || P > 256) ? 3 : 5; }
constexpr int OLD_W = 8;           // rolling window of prefetched x[t-dil] tiles (one per layer position)
constexpr int OLD_LA = 6;          // ... issued this many layer positions ahead
// fragment-ready block: [n-tile][k-step][lane][2 x u32]  (256 B per (n-tile, k-step))
constexpr int CONV_BYTES = 8 * 4 * 256;   // N = 64 (signal | gate), K = 64 (x[t-dil] | x[t])
constexpr int RES_BYTES = 4 * 2 * 256;    // N = 32, K = 32
constexpr int LAYER_A_BYTES = CONV_BYTES + RES_BYTES + 512;  // + biases: sig[32] gate[32] res[32] fp32 (pad to 512)
constexpr int chunk_bytes(int N) { return (N / 8) * 2 * 256; }  // [N / 8 n-tiles][2 k-steps (K = 32)]: SKIP_l, or a K slice of POST1 / POST2
constexpr int XP = 40;            // padded row (bf16 elements) of the 32-wide activation tiles
constexpr int hp(int S, int P) { return (S > P ? S : P) + 8; }   // padded row of the S / P-wide activation tiles
constexpr int LP = 264;           // padded row (floats) of the logits tile: rows g, g + 1 land in different banks
}  // namespace g2

struct Gen2Layout {  // byte offsets inside the generation-2 weight blob (device memory)
  int64_t layer_a;   // L x LAYER_A_BYTES
  int64_t skip;      // L x chunk_bytes(S)
  int64_t post1;     // S / 32 x chunk_bytes(P)
  int64_t post2;     // P / 32 x chunk_bytes(Q)
  int64_t x0tab;     // bf16 [257][32]: bf16(PRE[code] + bias), row 256 = all-zero input
  int64_t biases;    // fp32: skip-bias sum [S] | POST1_BIAS [P] | POST2_BIAS [Q]
  int64_t total;
};

static Gen2Layout gen2_layout(int L, int S, int P) {
  Gen2Layout g;
  int64_t off = 0;
  auto take = [&](int64_t b) { int64_t o = off; off = align_up(off + b, 1024); return o; };
  g.layer_a = take((int64_t)L * g2::LAYER_A_BYTES);
  g.skip = take((int64_t)L * g2::chunk_bytes(S));
  g.post1 = take((int64_t)(S / 32) * g2::chunk_bytes(P));
  g.post2 = take((int64_t)(P / 32) * g2::chunk_bytes(g2::Q));
  g.x0tab = take(257 * 32 * 2);
  g.biases = take((int64_t)(S + P + g2::Q) * 4);
  g.total = off;
  return g;
}

// value of B[k][n] for mma fragment register `reg` of `lane`: k = k0 + 2*(lane%4) + 8*reg (+0,+1), n = n0 + lane/4
__device__ __forceinline__ uint32_t frag_pack(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// Builds the fragment-ready blob from the fp32 arena.  One block per (kind, index).
__global__ void k_gen2_prep(const float* __restrict__ p, const LayerDesc* __restrict__ layers, int L, int64_t off_pre,
                            int64_t off_pre_b, int64_t off_post1, int64_t off_post1_b, int64_t off_post2,
                            int64_t off_post2_b, unsigned char* __restrict__ blob, Gen2Layout g, int S, int P) {
  using namespace g2;
  const int bid = blockIdx.x, tid = threadIdx.x;
  auto frag_store = [&](uint32_t* dst, int nt, int ks, int nks, auto Bfun) {
    // dst: [nt][ks][lane][2]
    for (int i = tid; i < nt * nks * 32; i += blockDim.x) {
      const int lane = i & 31, k_s = (i >> 5) % nks, n_t = (i >> 5) / nks;
      const int n = n_t * 8 + (lane >> 2), k = k_s * 16 + 2 * (lane & 3);
      dst[(size_t)i * 2 + 0] = frag_pack(Bfun(k, n), Bfun(k + 1, n));
      dst[(size_t)i * 2 + 1] = frag_pack(Bfun(k + 8, n), Bfun(k + 9, n));
    }
    (void)ks;
  };
  if (bid < L) {
    const LayerDesc ld = layers[bid];
    unsigned char* la = blob + g.layer_a + (size_t)bid * LAYER_A_BYTES;
    // conv: K index 0..31 = tap 0 (x[t-dil]), 32..63 = tap 1 (x[t]); N index 0..31 signal, 32..63 gate
    frag_store(reinterpret_cast<uint32_t*>(la), 8, 0, 4, [&](int k, int n) {
      const int tap = k >> 5, r = k & 31;
      return p[(n < D ? ld.sig : ld.gate) + ((int64_t)tap * R + r) * D + (n & 31)];
    });
    frag_store(reinterpret_cast<uint32_t*>(la + CONV_BYTES), 4, 0, 2,
               [&](int k, int n) { return p[ld.res + (int64_t)k * R + n]; });
    float* bias = reinterpret_cast<float*>(la + CONV_BYTES + RES_BYTES);
    for (int i = tid; i < 128; i += blockDim.x) {
      float v = 0.f;
      if (i < 32 && ld.sig_b >= 0) v = p[ld.sig_b + i];
      else if (i >= 32 && i < 64 && ld.gate_b >= 0) v = p[ld.gate_b + i - 32];
      else if (i >= 64 && i < 96 && ld.res_b >= 0) v = p[ld.res_b + i - 64];
      bias[i] = v;
    }
    frag_store(reinterpret_cast<uint32_t*>(blob + g.skip + (size_t)bid * chunk_bytes(S)), S / 8, 0, 2,
               [&](int k, int n) { return p[ld.skip + (int64_t)k * S + n]; });
  } else if (bid < L + S / 32) {
    const int c = bid - L;  // K slice [32c, 32c+32) of POST1 [S][P]
    frag_store(reinterpret_cast<uint32_t*>(blob + g.post1 + (size_t)c * chunk_bytes(P)), P / 8, 0, 2,
               [&](int k, int n) { return p[off_post1 + (int64_t)(c * 32 + k) * P + n]; });
  } else if (bid < L + S / 32 + P / 32) {
    const int c = bid - L - S / 32;
    frag_store(reinterpret_cast<uint32_t*>(blob + g.post2 + (size_t)c * chunk_bytes(Q)), Q / 8, 0, 2,
               [&](int k, int n) { return p[off_post2 + (int64_t)(c * 32 + k) * Q + n]; });
  } else {
    bf16* tab = reinterpret_cast<bf16*>(blob + g.x0tab);
    for (int i = tid; i < 257 * 32; i += blockDim.x) {
      const int code = i >> 5, r = i & 31;
      float v = code < 256 ? p[off_pre + (int64_t)code * R + r] : 0.f;
      if (off_pre_b >= 0) v += p[off_pre_b + r];
      tab[i] = f2bf(v);
    }
    float* b = reinterpret_cast<float*>(blob + g.biases);
    for (int i = tid; i < S + P + Q; i += blockDim.x) {
      float v = 0.f;
      const int which = i < S ? 0 : i < S + P ? 1 : 2, c = i < S ? i : i < S + P ? i - S : i - S - P;
      if (which == 0) {
        for (int l = 0; l < L; ++l)
          if (layers[l].skip_b >= 0) v += p[layers[l].skip_b + c];
      } else if (which == 1) {
        if (off_post1_b >= 0) v = p[off_post1_b + c];
      } else {
        if (off_post2_b >= 0) v = p[off_post2_b + c];
      }
      b[i] = v;
    }
  }
}

// ---- device helpers --------------------------------------------------------------------------------
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// A fragment (16 x 16) of a row-major bf16 tile in shared memory (row pitch `ld` elements), columns k0..k0+15
__device__ __forceinline__ void lda_frag(uint32_t (&a)[4], const bf16* tile, int ld, int k0, int lane) {
  const int g = lane >> 2, t = lane & 3;
  a[0] = *reinterpret_cast<const uint32_t*>(tile + g * ld + k0 + 2 * t);
  a[1] = *reinterpret_cast<const uint32_t*>(tile + (g + 8) * ld + k0 + 2 * t);
  a[2] = *reinterpret_cast<const uint32_t*>(tile + g * ld + k0 + 8 + 2 * t);
  a[3] = *reinterpret_cast<const uint32_t*>(tile + (g + 8) * ld + k0 + 8 + 2 * t);
}
__device__ __forceinline__ uint2 ldb_frag(const unsigned char* blk, int nt, int ks, int nks, int lane) {
  return *reinterpret_cast<const uint2*>(blk + ((size_t)(nt * nks + ks) * 32 + lane) * 8);
}
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 2, 256;" ::: "memory"); }  // compute warps only
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_pending() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_pending_dyn(int n) {  // at most n (0..5) most recent groups still pending
  switch (n) {
    case 0: cp_async_wait_pending<0>(); break;
    case 1: cp_async_wait_pending<1>(); break;
    case 2: cp_async_wait_pending<2>(); break;
    case 3: cp_async_wait_pending<3>(); break;
    case 4: cp_async_wait_pending<4>(); break;
    default: cp_async_wait_pending<5>(); break;
  }
}

struct Gen2Args {
  const unsigned char* blob;
  Gen2Layout g;
  const LayerDesc* layers;
  const int64_t* ring_off;
  bf16* rings;
  int32_t* codes;
  const int32_t* teacher;
  int32_t* out;
  float* logits_out;
  int64_t t0;
  uint64_t seed;
  int n_streams, n_steps, n_teacher, L;
  const float* gcproj;  // global conditioning: fp32 [n_streams][L][2D] per-stream projections (k_gen_gcproj), else nullptr
  long long* trace;  // wn_debug_trace buffer: CTA 0's warps log (event, layer, clock64) during the last step
};

template <int S, int P, bool GC>
__global__ void __launch_bounds__(g2::THREADS, 1) k_gen2(Gen2Args a) {
  using namespace g2;
  constexpr int SLOT = slot_bytes(S, P), NSLOT = n_slots(S, P), HP = hp(S, P);
  constexpr int NHS = S / 256, NHP = P / 256;   // 256-column halves: warp w owns n-tiles h * 32 + 4w .. 4w + 3 of every half
  constexpr int SKIP_BYTES = chunk_bytes(S), P1_BYTES = chunk_bytes(P), P2_BYTES = chunk_bytes(Q);
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* wring = sm;                                         // NSLOT x 27 KB
  bf16* x0tab = reinterpret_cast<bf16*>(sm + NSLOT * SLOT);          // [257][32]
  bf16* xbuf = x0tab + 257 * 32;                                     // [2][GS][XP]   current layer input (ping-pong)
  bf16* zbuf = xbuf + 2 * GS * XP;                                   // [GS][XP]
  bf16* hbuf = zbuf + GS * XP;                                       // [2][GS][HP]   h1 / h2
  float* lgbuf = reinterpret_cast<float*>(hbuf + 2 * GS * HP);       // [GS][LP]
  bf16* oldbuf = reinterpret_cast<bf16*>(lgbuf + GS * LP);            // [OLD_W][GS][XP] prefetched x[t-dil] tiles
  float* bias3 = reinterpret_cast<float*>(oldbuf + OLD_W * GS * XP);  // [S + P + Q]
  __shared__ __align__(8) uint64_t full[NSLOT], empty[NSLOT];
  __shared__ int code_s[GS];
  __shared__ int dil_s[64];
  __shared__ int64_t roff_s[64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s0 = blockIdx.x * GS;
  const int L = a.L;

  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], NCW);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < 257 * 32 / 8; i += THREADS)
    reinterpret_cast<uint4*>(x0tab)[i] = reinterpret_cast<const uint4*>(a.blob + a.g.x0tab)[i];
  for (int i = tid; i < S + P + Q; i += THREADS) bias3[i] = reinterpret_cast<const float*>(a.blob + a.g.biases)[i];
  if (tid < GS) code_s[tid] = (s0 + tid < a.n_streams) ? a.codes[s0 + tid] : -1;
  if (tid < L) {
    dil_s[tid] = a.layers[tid].dil;
    roff_s[tid] = a.ring_off[tid];
  }
  __syncthreads();

  const int items_per_step = L + S / 32 + P / 32;
  if (warp == NCW) {
    // ===== weight producer: same item sequence every timestep =====
    // One item per layer (conv | residual | biases, then SKIP_l: two bulk copies into one slot) and 16 post-net chunks.
    // (The in-kernel timeline showed ~290 cycles per wait / release pair on the per-sample chain: one pair per layer.)
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 1;
      for (int step = 0; step < a.n_steps; ++step) {
        for (int j = 0; j < items_per_step; ++j) {
          mbar_wait(&empty[slot], ph);
          unsigned char* dst = wring + slot * SLOT;
          if (j < L) {
            mbar_expect_tx(&full[slot], LAYER_A_BYTES + SKIP_BYTES);
            bulk_g2s(dst, a.blob + a.g.layer_a + (size_t)j * LAYER_A_BYTES, LAYER_A_BYTES, &full[slot]);
            bulk_g2s(dst + LAYER_A_BYTES, a.blob + a.g.skip + (size_t)j * SKIP_BYTES, SKIP_BYTES, &full[slot]);
          } else if (j < L + S / 32) {
            mbar_expect_tx(&full[slot], P1_BYTES);
            bulk_g2s(dst, a.blob + a.g.post1 + (size_t)(j - L) * P1_BYTES, P1_BYTES, &full[slot]);
          } else {
            mbar_expect_tx(&full[slot], P2_BYTES);
            bulk_g2s(dst, a.blob + a.g.post2 + (size_t)(j - L - S / 32) * P2_BYTES, P2_BYTES, &full[slot]);
          }
          if (++slot == NSLOT) { slot = 0; ph ^= 1u; }
        }
      }
    }
    return;
  }

  // ===== compute warps =====
  Tracer tr;
  tr.init(nullptr, warp, false);
  const int g = lane >> 2, t4 = lane & 3;
  // weight-ring consumer positions (same item sequence as the producer).  Waits and releases advance separately:
  // a layer's slot is held until its skip MMAs have run, which is during the NEXT layer (below)
  int w_slot = 0, r_slot = 0;
  uint32_t w_ph = 0;
  auto slot_wait = [&]() -> const unsigned char* {
    mbar_wait(&full[w_slot], w_ph);
    const unsigned char* p = wring + w_slot * SLOT;
    if (++w_slot == NSLOT) { w_slot = 0; w_ph ^= 1u; }
    return p;
  };
  auto slot_release = [&]() {
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[r_slot]);
    if (++r_slot == NSLOT) r_slot = 0;
  };
  // x[t-dil] of every layer comes from its HBM ring buffer (length dil, slot t mod dil; dil = 2^k, dil == 1 included:
  // the slot then holds the previous step's input).  Warps 4..7 prefetch it with cp.async `la` layer positions ahead
  // (position p = step * L + l; the addresses depend on p only) into a rolling window of OLD_W tiles.  x_l[t] is
  // stored at position (t, l) and read for (t + dil, l), whose prefetch is issued at (t + dil, l) - la > (t, l)
  // because la = min(OLD_LA, L - 1) < L.  (A whole step ahead, as before, cost 77 KB of shared memory.)
  const int la = max(1, min(OLD_LA, L - 1));
  int pf_l = 0, pf_buf = 0;
  int64_t pf_t = a.t0, pf_left = (int64_t)a.n_steps * L;
  auto prefetch_next = [&]() {  // warps 4..7 only
    if (pf_left > 0) {
      const int q = tid - 4 * 32;
      if (q < GS * 4) {
        const int s = q >> 2, ch = q & 3, dil = dil_s[pf_l];
        if (s0 + s < a.n_streams)
          cp_async16(oldbuf + ((size_t)pf_buf * GS + s) * XP + ch * 8,
                     a.rings + roff_s[pf_l] + ((int64_t)(s0 + s) * dil + (pf_t & (int64_t)(dil - 1))) * R + ch * 8);
      }
      if (++pf_l == L) { pf_l = 0; ++pf_t; }
      pf_buf = (pf_buf + 1) & (OLD_W - 1);
      --pf_left;
    }
    cp_async_commit();
  };
  if (warp >= 4) {
    for (int k = 0; k < la; ++k) prefetch_next();
    cp_async_wait_pending_dyn(la - 1);  // position 0 has landed
  }
  cbar();
  int cur_buf = 0;  // window tile of the current layer position

  for (int step = 0; step < a.n_steps; ++step) {
    const int64_t t = a.t0 + step;
    if (step == a.n_steps - 1) tr.init(a.trace, warp, blockIdx.x == 0 && lane == 0);
    tr.ev(20, 0);
    // the sampler's uniforms depend on (seed, t, stream) only: computed here, off the tail of the step
    const float ua = sampler_uniform(a.seed, (uint64_t)t, (uint32_t)(s0 + warp * 2));
    const float ub = sampler_uniform(a.seed, (uint64_t)t, (uint32_t)(s0 + warp * 2 + 1));
    // input embedding (imodel.py:66-74): table row of the pending code; -1 -> all-zero vector -> bias only
    for (int i = tid; i < GS * 4; i += NCW * 32) {
      const int s = i >> 2, ch = i & 3;
      const int c = code_s[s];
      const int row = (c >= 0 && c < 256) ? c : 256;
      *reinterpret_cast<uint4*>(xbuf + s * XP + ch * 8) = *reinterpret_cast<const uint4*>(x0tab + row * 32 + ch * 8);
    }
    float skip[NHS * 4][4];
#pragma unroll
    for (int i = 0; i < NHS * 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) skip[i][j] = 0.f;
    // The skip MMAs of layer l (A = z_l, off the per-sample chain) are deferred into layer l + 1, where they fill the
    // latency bubbles of the conv chain instead of sitting between the residual and the second barrier
    uint32_t zf[2][4];                         // A fragments of the previous layer's z
    const unsigned char* wsk_prev = nullptr;   // its SKIP weights (slot still held)
    auto skip_mma = [&]() {
#pragma unroll
      for (int h = 0; h < NHS; ++h) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint2 b = ldb_frag(wsk_prev, h * 32 + warp * 4 + j, ks, 2, lane);
            mma16816(skip[h * 4 + j], zf[ks], b.x, b.y);  // imodel.py:247 (bias added once, below)
          }
        }
      }
      slot_release();
    };
    // global conditioning: this lane's per-stream projections of the CURRENT layer (rows g, g + 8; channels c, c + 1 of
    // the signal and the gate half), loaded one layer ahead -- they depend on (stream, layer) only
    float2 gcs[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)}, gcg[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    float2 ngs[2], ngg[2];
    auto gc_load = [&](int l) {
      if constexpr (GC) {
        if (warp < 4) {
          const int c = warp * 8 + 2 * (lane & 3);
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int sidx = min(s0 + (lane >> 2) + 8 * k, a.n_streams - 1);
            const float* q = a.gcproj + ((size_t)sidx * L + l) * (2 * D) + c;
            ngs[k] = __ldg(reinterpret_cast<const float2*>(q));
            ngg[k] = __ldg(reinterpret_cast<const float2*>(q + D));
          }
        }
      }
    };
    gc_load(0);
    cbar();
    int xb = 0;
    for (int l = 0; l < L; ++l) {
      const int dil = dil_s[l];
      const bf16* xin = xbuf + xb * GS * XP;
      bf16* xout = xbuf + (xb ^ 1) * GS * XP;
      const bf16* oldx = oldbuf + cur_buf * (GS * XP);
      tr.ev(0, l);
      if constexpr (GC) {
        gcs[0] = ngs[0]; gcs[1] = ngs[1]; gcg[0] = ngg[0]; gcg[1] = ngg[1];
        if (l + 1 < L) gc_load(l + 1);
      }
      const unsigned char* wa = slot_wait();
      tr.ev(1, l);
      const float* bias = reinterpret_cast<const float*>(wa + CONV_BYTES + RES_BYTES);
      if (warp < 4) {
        // conv + gate: this warp owns channels [8w, 8w+8): signal n-tile w, gate n-tile w + 4
        // (the x[t-dil] and x[t] halves of K accumulate separately: two dependent MMAs deep instead of four)
        float cs[4] = {0.f, 0.f, 0.f, 0.f}, cg[4] = {0.f, 0.f, 0.f, 0.f};
        float ds[4] = {0.f, 0.f, 0.f, 0.f}, dg[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t af[4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          lda_frag(af, ks < 2 ? oldx : xin, XP, (ks & 1) * 16, lane);
          const uint2 bs = ldb_frag(wa, warp, ks, 4, lane), bg = ldb_frag(wa, warp + 4, ks, 4, lane);
          if (ks < 2) {
            mma16816(cs, af, bs.x, bs.y);
            mma16816(cg, af, bg.x, bg.y);
          } else {
            mma16816(ds, af, bs.x, bs.y);
            mma16816(dg, af, bg.x, bg.y);
          }
        }
        if (l > 0) skip_mma();  // the previous layer's skip contribution, in the shadow of the conv chain
        tr.ev(2, l);
        const int c = warp * 8 + 2 * t4;
        const float bs0 = bias[c], bs1 = bias[c + 1], bg0 = bias[32 + c], bg1 = bias[32 + c + 1];
        // (+ the stream's global-conditioning projections, imodel.py:113-118; zero without GC)
        const float z00 = tanh_fast(cs[0] + ds[0] + bs0 + gcs[0].x) * sigmoid_fast(cg[0] + dg[0] + bg0 + gcg[0].x);
        const float z01 = tanh_fast(cs[1] + ds[1] + bs1 + gcs[0].y) * sigmoid_fast(cg[1] + dg[1] + bg1 + gcg[0].y);
        const float z10 = tanh_fast(cs[2] + ds[2] + bs0 + gcs[1].x) * sigmoid_fast(cg[2] + dg[2] + bg0 + gcg[1].x);
        const float z11 = tanh_fast(cs[3] + ds[3] + bs1 + gcs[1].y) * sigmoid_fast(cg[3] + dg[3] + bg1 + gcg[1].y);
        *reinterpret_cast<uint32_t*>(zbuf + g * XP + c) = frag_pack(z00, z01);
        *reinterpret_cast<uint32_t*>(zbuf + (g + 8) * XP + c) = frag_pack(z10, z11);
      } else {
        // meanwhile: ring <- x[t] (imodel.py:97), slot t mod dil; then the read of a later position
        const int i = (warp - 4) * 32 + lane;  // 128 threads: 16 streams x 4 chunks of 16 B
        if (i < GS * 4) {
          const int s = i >> 2, ch = i & 3;
          const uint4 v = *reinterpret_cast<const uint4*>(xin + s * XP + ch * 8);
          if (s0 + s < a.n_streams)
            *reinterpret_cast<uint4*>(a.rings + roff_s[l] + ((int64_t)(s0 + s) * dil + (t & (int64_t)(dil - 1))) * R + ch * 8) = v;
        }
        prefetch_next();
        if (l > 0) skip_mma();
        cp_async_wait_pending_dyn(la - 1);  // the next position's tile has landed (visible after the barriers below)
      }
      tr.ev(3, l);
      cbar();
      tr.ev(4, l);
      {
        // residual (warps 0..3: n-tile w), A = z; the z fragments stay in registers for the deferred skip MMAs
        lda_frag(zf[0], zbuf, XP, 0, lane);
        lda_frag(zf[1], zbuf, XP, 16, lane);
        if (warp < 4 && l + 1 < L) {
          float cr[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint2 b = ldb_frag(wa + CONV_BYTES, warp, ks, 2, lane);
            mma16816(cr, zf[ks], b.x, b.y);
          }
          const int c = warp * 8 + 2 * t4;
          const float br0 = bias[64 + c], br1 = bias[64 + c + 1];
          const __nv_bfloat162 x_lo = *reinterpret_cast<const __nv_bfloat162*>(xin + g * XP + c);
          const __nv_bfloat162 x_hi = *reinterpret_cast<const __nv_bfloat162*>(xin + (g + 8) * XP + c);
          *reinterpret_cast<uint32_t*>(xout + g * XP + c) =
              frag_pack(__low2float(x_lo) + cr[0] + br0, __high2float(x_lo) + cr[1] + br1);  // imodel.py:245
          *reinterpret_cast<uint32_t*>(xout + (g + 8) * XP + c) =
              frag_pack(__low2float(x_hi) + cr[2] + br0, __high2float(x_hi) + cr[3] + br1);
        }
        tr.ev(5, l);
        wsk_prev = wa + LAYER_A_BYTES;
      }
      cbar();
      tr.ev(8, l);
      xb ^= 1;
      cur_buf = (cur_buf + 1) & (OLD_W - 1);
    }
    skip_mma();  // the last layer's skip contribution
    tr.ev(9, 0);
    // ---- post-net (imodel.py:140-164) ----
    bf16* h1 = hbuf;
    bf16* h2 = hbuf + GS * HP;
#pragma unroll
    for (int hj = 0; hj < NHS * 4; ++hj) {
      const int c = ((hj >> 2) * 32 + warp * 4 + (hj & 3)) * 8 + 2 * t4;
      *reinterpret_cast<uint32_t*>(h1 + g * HP + c) =
          frag_pack(fmaxf(skip[hj][0] + bias3[c], 0.f), fmaxf(skip[hj][1] + bias3[c + 1], 0.f));
      *reinterpret_cast<uint32_t*>(h1 + (g + 8) * HP + c) =
          frag_pack(fmaxf(skip[hj][2] + bias3[c], 0.f), fmaxf(skip[hj][3] + bias3[c + 1], 0.f));
    }
    cbar();
    // acc[h * 4 + j] = A[16 x K] . W[K x (n-tiles h * 32 + 4w .. 4w + 3)], W streamed in K / 32 slices of [32 x N]
    float acc[(NHP > 1 ? NHP : 1) * 4][4];
    auto dense = [&](const bf16* A, auto kc, auto nh) {
      constexpr int KC = decltype(kc)::value, NH = decltype(nh)::value;
#pragma unroll
      for (int i = 0; i < NH * 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int c8 = 0; c8 < KC; ++c8) {
        const unsigned char* w = slot_wait();
        uint32_t af[4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          lda_frag(af, A, HP, c8 * 32 + ks * 16, lane);
#pragma unroll
          for (int h = 0; h < NH; ++h) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint2 b = ldb_frag(w, h * 32 + warp * 4 + j, ks, 2, lane);
              mma16816(acc[h * 4 + j], af, b.x, b.y);
            }
          }
        }
        slot_release();
      }
    };
    dense(h1, std::integral_constant<int, S / 32>{}, std::integral_constant<int, NHP>{});
#pragma unroll
    for (int hj = 0; hj < NHP * 4; ++hj) {
      const int c = ((hj >> 2) * 32 + warp * 4 + (hj & 3)) * 8 + 2 * t4;
      *reinterpret_cast<uint32_t*>(h2 + g * HP + c) =
          frag_pack(fmaxf(acc[hj][0] + bias3[S + c], 0.f), fmaxf(acc[hj][1] + bias3[S + c + 1], 0.f));
      *reinterpret_cast<uint32_t*>(h2 + (g + 8) * HP + c) =
          frag_pack(fmaxf(acc[hj][2] + bias3[S + c], 0.f), fmaxf(acc[hj][3] + bias3[S + c + 1], 0.f));
    }
    tr.ev(10, 0);
    cbar();
    dense(h2, std::integral_constant<int, P / 32>{}, std::integral_constant<int, 1>{});
    tr.ev(11, 0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (warp * 4 + j) * 8 + 2 * t4;
      *reinterpret_cast<float2*>(lgbuf + g * LP + c) = make_float2(acc[j][0] + bias3[S + P + c], acc[j][1] + bias3[S + P + c + 1]);
      *reinterpret_cast<float2*>(lgbuf + (g + 8) * LP + c) = make_float2(acc[j][2] + bias3[S + P + c], acc[j][3] + bias3[S + P + c + 1]);
    }
    cbar();
    // ---- sampling (imodel.py:167-187) + teacher forcing (imodel.py:260-267): warp w -> streams 2w, 2w+1, in lock step ----
    {
      const int sa = warp * 2, sb = warp * 2 + 1;
      if (a.logits_out != nullptr) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
          if (s0 + sa + k < a.n_streams)
            for (int q = lane; q < Q; q += 32)
              a.logits_out[((int64_t)(s0 + sa + k) * a.n_steps + step) * Q + q] = lgbuf[(sa + k) * LP + q];
      }
      int ra, rb;
      warp_sample2(lgbuf + sa * LP, lgbuf + sb * LP, ua, ub, ra, rb);
      if (lane < 2) {
        const int s = sa + lane, samp = lane == 0 ? ra : rb;
        if (s0 + s < a.n_streams) {
          a.out[(int64_t)(s0 + s) * a.n_steps + step] = samp;
          code_s[s] = (t < a.n_teacher) ? a.teacher[t] : samp;
        }
      }
    }
    tr.ev(12, 0);
    cbar();
    tr.ev(13, 0);
  }
  // persist the state a later launch continues from: the pending codes (the rings already hold every x[t-dil])
  if (tid < GS && s0 + tid < a.n_streams) a.codes[s0 + tid] = code_s[tid];
}

// =====================================================================================================
// Generation 3: the per-sample chain on ONE warp, no CTA barrier inside the layer loop.
//
// k_gen2 spreads a layer over 8 warps and pays for it on the dependency chain: two 256-thread barriers per layer plus
// the shared-memory round trips of z and x (~700 of ~1400 cycles per layer, in-kernel timeline of round 1).  Here warp 0
// computes a whole layer for its 16 streams by itself -- 32 mma.sync for the conv (all 64 SIGNAL | GATE columns), the
// gate, 8 mma.sync for the residual -- and hands x from layer to layer IN REGISTERS: the accumulator fragment of an
// m16n8k16 (rows g, g+8; columns 2t, 2t+1 of an 8-column n-tile) is exactly half of the A fragment of the next
// contraction, so z feeds the residual and x' feeds the next conv without touching shared memory.  Measured
// (tools/hmma_cost.cu): ~680 cycles per layer for that arithmetic on one warp.  Everything else is taken off the chain:
//   warp 0          the chain; publishes z_l and x_l (two 1.3 KB tiles) through mbarrier-guarded double buffers
//   warps 1,2,3,5,6,7  skip contraction z_l . SKIP_l (each owns 1/6 of the n-tiles; accumulators in registers), on the
//                   three schedulers the chain warp does not run on
//   warp 4          ring buffers: x_l[t] -> HBM, cp.async prefetch of x[t-dil] six layer positions ahead
//   warp 8          weight producer (cp.async.bulk into the slot ring), as in k_gen2
// The post-net and the sampler are k_gen2's: all eight compute warps, three barriers per TIMESTEP.
// =====================================================================================================
template <int S, int P, bool GC>
__global__ void __launch_bounds__(g2::THREADS, 1) k_gen3(Gen2Args a) {
  using namespace g2;
  constexpr int SLOT = slot_bytes(S, P), NSLOT = n_slots(S, P), HP = hp(S, P);
  constexpr int NHP = P / 256;
  constexpr int SKIP_BYTES = chunk_bytes(S), P1_BYTES = chunk_bytes(P), P2_BYTES = chunk_bytes(Q);
  constexpr int NTS = S / 8, NHW = 6, CNT = (NTS + NHW - 1) / NHW;  // skip n-tiles, helper warps, n-tiles per helper
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned char* wring = sm;                                         // NSLOT x SLOT
  bf16* x0tab = reinterpret_cast<bf16*>(sm + NSLOT * SLOT);          // [257][32]
  bf16* xst = x0tab + 257 * 32;                                      // [2][GS][XP]   x_l published for the ring writer
  bf16* zbuf = xst + 2 * GS * XP;                                    // [2][GS][XP]   z_l published for the skip warps
  bf16* hbuf = zbuf + 2 * GS * XP;                                   // [2][GS][HP]   h1 / h2
  float* lgbuf = reinterpret_cast<float*>(hbuf + 2 * GS * HP);       // [GS][LP]
  bf16* oldbuf = reinterpret_cast<bf16*>(lgbuf + GS * LP);            // [OLD_W][GS][XP] prefetched x[t-dil] tiles
  float* bias3 = reinterpret_cast<float*>(oldbuf + OLD_W * GS * XP);  // [S + P + Q]
  __shared__ __align__(8) uint64_t full[NSLOT], empty[NSLOT], z_ready[2], z_free[2], x_ready[2], x_free[2], old_ready[OLD_W];
  __shared__ int code_s[GS];
  __shared__ int dil_s[64];
  __shared__ int64_t roff_s[64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int s0 = blockIdx.x * GS;
  const int L = a.L;

  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], NCW);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&z_ready[i], 1);
      mbar_init(&z_free[i], NHW);
      mbar_init(&x_ready[i], 1);
      mbar_init(&x_free[i], 1);
    }
    for (int i = 0; i < OLD_W; ++i) mbar_init(&old_ready[i], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < 257 * 32 / 8; i += THREADS)
    reinterpret_cast<uint4*>(x0tab)[i] = reinterpret_cast<const uint4*>(a.blob + a.g.x0tab)[i];
  for (int i = tid; i < S + P + Q; i += THREADS) bias3[i] = reinterpret_cast<const float*>(a.blob + a.g.biases)[i];
  if (tid < GS) code_s[tid] = (s0 + tid < a.n_streams) ? a.codes[s0 + tid] : -1;
  if (tid < L) {
    dil_s[tid] = a.layers[tid].dil;
    roff_s[tid] = a.ring_off[tid];
  }
  __syncthreads();

  const int items_per_step = L + S / 32 + P / 32;
  if (warp == NCW) {  // ===== weight producer: the same item sequence every timestep (as k_gen2) =====
    if (lane == 0) {
      int slot = 0;
      uint32_t ph = 1;
      for (int step = 0; step < a.n_steps; ++step) {
        for (int j = 0; j < items_per_step; ++j) {
          mbar_wait(&empty[slot], ph);
          unsigned char* dst = wring + slot * SLOT;
          if (j < L) {
            mbar_expect_tx(&full[slot], LAYER_A_BYTES + SKIP_BYTES);
            bulk_g2s(dst, a.blob + a.g.layer_a + (size_t)j * LAYER_A_BYTES, LAYER_A_BYTES, &full[slot]);
            bulk_g2s(dst + LAYER_A_BYTES, a.blob + a.g.skip + (size_t)j * SKIP_BYTES, SKIP_BYTES, &full[slot]);
          } else if (j < L + S / 32) {
            mbar_expect_tx(&full[slot], P1_BYTES);
            bulk_g2s(dst, a.blob + a.g.post1 + (size_t)(j - L) * P1_BYTES, P1_BYTES, &full[slot]);
          } else {
            mbar_expect_tx(&full[slot], P2_BYTES);
            bulk_g2s(dst, a.blob + a.g.post2 + (size_t)(j - L - S / 32) * P2_BYTES, P2_BYTES, &full[slot]);
          }
          if (++slot == NSLOT) { slot = 0; ph ^= 1u; }
        }
      }
    }
    return;
  }

  // ===== compute warps =====
  const int g = lane >> 2, t4 = lane & 3;
  int w_slot = 0, r_slot = 0;
  uint32_t w_ph = 0;
  auto slot_wait = [&]() -> const unsigned char* {
    mbar_wait(&full[w_slot], w_ph);
    const unsigned char* p = wring + w_slot * SLOT;
    if (++w_slot == NSLOT) { w_slot = 0; w_ph ^= 1u; }
    return p;
  };
  auto slot_release = [&]() {
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[r_slot]);
    if (++r_slot == NSLOT) r_slot = 0;
  };
  auto publish = [&](uint64_t* bar) {  // this warp's shared-memory writes / reads are done: one arrival
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
  };
  const bool is_chain = warp == 0, is_io = warp == 4;
  const int hw = warp < 4 ? warp - 1 : warp - 2;                  // skip helper index 0..5 (warps 1,2,3,5,6,7)
  const int nt0 = is_chain || is_io ? 0 : (hw * NTS) / NHW;       // this helper's skip n-tiles [nt0, nt1)
  const int nt1 = is_chain || is_io ? 0 : ((hw + 1) * NTS) / NHW;

  // ---- ring-buffer warp state: x[t-dil] of layer position p = step * L + l is prefetched la positions ahead ----
  const int la = max(1, min(OLD_LA, L - 1));
  int pf_l = 0, pf_buf = 0;
  int64_t pf_t = a.t0, pf_left = (int64_t)a.n_steps * L;
  auto prefetch_next = [&]() {  // warp 4: 64 chunks of 16 bytes, two per lane
    if (pf_left > 0) {
      const int dil = dil_s[pf_l];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int q = lane + 32 * u, s = q >> 2, ch = q & 3;
        if (s0 + s < a.n_streams)
          cp_async16(oldbuf + ((size_t)pf_buf * GS + s) * XP + ch * 8,
                     a.rings + roff_s[pf_l] + ((int64_t)(s0 + s) * dil + (pf_t & (int64_t)(dil - 1))) * R + ch * 8);
      }
      if (++pf_l == L) { pf_l = 0; ++pf_t; }
      pf_buf = (pf_buf + 1) & (OLD_W - 1);
      --pf_left;
    }
    cp_async_commit();
  };
  if (is_io) {
    for (int k = 0; k < la; ++k) prefetch_next();
    cp_async_wait_pending_dyn(la - 1);  // position 0 has landed
    publish(&old_ready[0]);
  }

  int64_t pos = 0;  // layer position since the start of this launch
  // developer aid (wn_debug_trace): cycles the chain warp of CTA 0 spends per section of a layer, summed over the launch
  const bool prof = a.trace != nullptr && blockIdx.x == 0 && warp == 0;
  long long pacc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, pt = 0;
#define G3_MARK(k_)                    \
  if (prof) {                          \
    const long long now_ = clock64(); \
    pacc[k_] += now_ - pt;             \
    pt = now_;                         \
  }
  for (int step = 0; step < a.n_steps; ++step) {
    const int64_t t = a.t0 + step;
    // the sampler's uniforms depend on (seed, t, stream) only: computed here, off the tail of the step
    const float ua = sampler_uniform(a.seed, (uint64_t)t, (uint32_t)(s0 + warp * 2));
    const float ub = sampler_uniform(a.seed, (uint64_t)t, (uint32_t)(s0 + warp * 2 + 1));
    float skip[CNT][4];
#pragma unroll
    for (int i = 0; i < CNT; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) skip[i][j] = 0.f;

    if (is_chain) {
      // ---- input embedding (imodel.py:66-74) straight into A fragments: rows g, g + 8 of the 16-stream tile ----
      uint32_t xa[2][4];
      {
        const int c0 = code_s[g], c1 = code_s[g + 8];
        const bf16* r0 = x0tab + ((c0 >= 0 && c0 < 256) ? c0 : 256) * 32;
        const bf16* r1 = x0tab + ((c1 >= 0 && c1 < 256) ? c1 : 256) * 32;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          xa[ks][0] = *reinterpret_cast<const uint32_t*>(r0 + ks * 16 + 2 * t4);
          xa[ks][1] = *reinterpret_cast<const uint32_t*>(r1 + ks * 16 + 2 * t4);
          xa[ks][2] = *reinterpret_cast<const uint32_t*>(r0 + ks * 16 + 8 + 2 * t4);
          xa[ks][3] = *reinterpret_cast<const uint32_t*>(r1 + ks * 16 + 8 + 2 * t4);
        }
      }
      // global conditioning: per-stream projections of the next layer, prefetched (rows g, g + 8; n-tile j: columns
      // 8j + 2t, +1 of the signal and of the gate half)
      float2 ngs[2][4], ngg[2][4];
      auto gc_load = [&](int l) {
        if constexpr (GC) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int sidx = min(s0 + g + 8 * k, a.n_streams - 1);
            const float* q = a.gcproj + ((size_t)sidx * L + l) * (2 * D) + 2 * t4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              ngs[k][j] = __ldg(reinterpret_cast<const float2*>(q + 8 * j));
              ngg[k][j] = __ldg(reinterpret_cast<const float2*>(q + D + 8 * j));
            }
          }
        }
      };
      gc_load(0);
      if (prof) pt = clock64();
      for (int l = 0; l < L; ++l, ++pos) {
        const int pb = (int)(pos & 1);
        const uint32_t ph_prev = (uint32_t)((pos - 2) >> 1) & 1u;
        // publish x_l[t] for the ring writer (imodel.py:97)
        bf16* xs = xst + pb * GS * XP;
        if (pos >= 2) mbar_wait(&x_free[pb], ph_prev);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          *reinterpret_cast<uint32_t*>(xs + g * XP + ks * 16 + 2 * t4) = xa[ks][0];
          *reinterpret_cast<uint32_t*>(xs + (g + 8) * XP + ks * 16 + 2 * t4) = xa[ks][1];
          *reinterpret_cast<uint32_t*>(xs + g * XP + ks * 16 + 8 + 2 * t4) = xa[ks][2];
          *reinterpret_cast<uint32_t*>(xs + (g + 8) * XP + ks * 16 + 8 + 2 * t4) = xa[ks][3];
        }
        publish(&x_ready[pb]);
        G3_MARK(0)
        float2 gcs[2][4], gcg[2][4];
        if constexpr (GC) {
#pragma unroll
          for (int k = 0; k < 2; ++k)
#pragma unroll
            for (int j = 0; j < 4; ++j) { gcs[k][j] = ngs[k][j]; gcg[k][j] = ngg[k][j]; }
          if (l + 1 < L) gc_load(l + 1);
        }
        const unsigned char* wa = slot_wait();
        G3_MARK(1)
        const float* bias = reinterpret_cast<const float*>(wa + CONV_BYTES + RES_BYTES);
        // x[t-dil] of this layer (prefetched tile)
        const int ob = (int)(pos & (OLD_W - 1));
        mbar_wait(&old_ready[ob], (uint32_t)(pos / OLD_W) & 1u);
        uint32_t oa[2][4];
        lda_frag(oa[0], oldbuf + ob * (GS * XP), XP, 0, lane);
        lda_frag(oa[1], oldbuf + ob * (GS * XP), XP, 16, lane);
        G3_MARK(2)
        // conv (imodel.py:107-108): n-tiles 0..3 = SIGNAL channels 8j.., 4..7 = GATE; K = x[t-dil] (32) | x[t] (32)
        float cv[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
          for (int j = 0; j < 4; ++j) cv[nt][j] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
            const uint2 b = ldb_frag(wa, nt, ks, 4, lane);
            if (ks < 2) mma16816(cv[nt], oa[ks], b.x, b.y);
            else mma16816(cv[nt], xa[ks - 2], b.x, b.y);
          }
        }
        G3_MARK(3)
        // gate (imodel.py:121) -> z as A fragments of the residual contraction, and as a tile for the skip warps
        uint32_t za[2][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = 8 * j + 2 * t4;
          const float2 bs = *reinterpret_cast<const float2*>(bias + c), bg = *reinterpret_cast<const float2*>(bias + 32 + c);
          float gs0 = 0.f, gs1 = 0.f, gs2 = 0.f, gs3 = 0.f, gg0 = 0.f, gg1 = 0.f, gg2 = 0.f, gg3 = 0.f;
          if constexpr (GC) {
            gs0 = gcs[0][j].x; gs1 = gcs[0][j].y; gs2 = gcs[1][j].x; gs3 = gcs[1][j].y;
            gg0 = gcg[0][j].x; gg1 = gcg[0][j].y; gg2 = gcg[1][j].x; gg3 = gcg[1][j].y;
          }
          const float z0 = tanh_fast(cv[j][0] + bs.x + gs0) * sigmoid_fast(cv[4 + j][0] + bg.x + gg0);
          const float z1 = tanh_fast(cv[j][1] + bs.y + gs1) * sigmoid_fast(cv[4 + j][1] + bg.y + gg1);
          const float z2 = tanh_fast(cv[j][2] + bs.x + gs2) * sigmoid_fast(cv[4 + j][2] + bg.x + gg2);
          const float z3 = tanh_fast(cv[j][3] + bs.y + gs3) * sigmoid_fast(cv[4 + j][3] + bg.y + gg3);
          za[j >> 1][(j & 1) * 2 + 0] = frag_pack(z0, z1);   // row g
          za[j >> 1][(j & 1) * 2 + 1] = frag_pack(z2, z3);   // row g + 8
        }
        G3_MARK(4)
        bf16* zs = zbuf + pb * GS * XP;
        if (pos >= 2) mbar_wait(&z_free[pb], ph_prev);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          *reinterpret_cast<uint32_t*>(zs + g * XP + ks * 16 + 2 * t4) = za[ks][0];
          *reinterpret_cast<uint32_t*>(zs + (g + 8) * XP + ks * 16 + 2 * t4) = za[ks][1];
          *reinterpret_cast<uint32_t*>(zs + g * XP + ks * 16 + 8 + 2 * t4) = za[ks][2];
          *reinterpret_cast<uint32_t*>(zs + (g + 8) * XP + ks * 16 + 8 + 2 * t4) = za[ks][3];
        }
        publish(&z_ready[pb]);
        G3_MARK(5)
        if (l + 1 < L) {
          // residual 1x1 + add (imodel.py:131,245): x' = bf16(x + z . RESIDUAL + b), straight back into A fragments
          float rr[4][4];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int j = 0; j < 4; ++j) rr[nt][j] = 0.f;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              const uint2 b = ldb_frag(wa + CONV_BYTES, nt, ks, 2, lane);
              mma16816(rr[nt], za[ks], b.x, b.y);
            }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 br = *reinterpret_cast<const float2*>(bias + 64 + 8 * j + 2 * t4);
            const uint32_t xlo = xa[j >> 1][(j & 1) * 2], xhi = xa[j >> 1][(j & 1) * 2 + 1];
            xa[j >> 1][(j & 1) * 2] = frag_pack(__uint_as_float(xlo << 16) + rr[j][0] + br.x,
                                                __uint_as_float(xlo & 0xffff0000u) + rr[j][1] + br.y);
            xa[j >> 1][(j & 1) * 2 + 1] = frag_pack(__uint_as_float(xhi << 16) + rr[j][2] + br.x,
                                                    __uint_as_float(xhi & 0xffff0000u) + rr[j][3] + br.y);
          }
        }
        G3_MARK(6)
        slot_release();
        G3_MARK(7)
      }
    } else if (is_io) {
      for (int l = 0; l < L; ++l, ++pos) {
        const int pb = (int)(pos & 1), dil = dil_s[l];
        (void)slot_wait();
        slot_release();
        // ring <- x_l[t] (imodel.py:97), slot t mod dil
        mbar_wait(&x_ready[pb], (uint32_t)(pos >> 1) & 1u);
        const bf16* xs = xst + pb * GS * XP;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int q = lane + 32 * u, s = q >> 2, ch = q & 3;
          const uint4 v = *reinterpret_cast<const uint4*>(xs + s * XP + ch * 8);
          if (s0 + s < a.n_streams)
            *reinterpret_cast<uint4*>(a.rings + roff_s[l] + ((int64_t)(s0 + s) * dil + (t & (int64_t)(dil - 1))) * R + ch * 8) = v;
        }
        publish(&x_free[pb]);
        prefetch_next();
        cp_async_wait_pending_dyn(la - 1);  // the next position's tile has landed
        if (pos + 1 < (int64_t)a.n_steps * L) publish(&old_ready[(pos + 1) & (OLD_W - 1)]);
      }
    } else {
      for (int l = 0; l < L; ++l, ++pos) {
        const int pb = (int)(pos & 1);
        const unsigned char* wsk = slot_wait() + LAYER_A_BYTES;
        mbar_wait(&z_ready[pb], (uint32_t)(pos >> 1) & 1u);
        uint32_t zf[2][4];
        lda_frag(zf[0], zbuf + pb * GS * XP, XP, 0, lane);
        lda_frag(zf[1], zbuf + pb * GS * XP, XP, 16, lane);
        publish(&z_free[pb]);
#pragma unroll
        for (int j = 0; j < CNT; ++j) {
          if (nt0 + j < nt1) {
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint2 b = ldb_frag(wsk, nt0 + j, ks, 2, lane);
              mma16816(skip[j], zf[ks], b.x, b.y);  // imodel.py:247 (bias added once, below)
            }
          }
        }
        slot_release();
      }
    }
    G3_MARK(8)
    // ---- post-net (imodel.py:140-164): as k_gen2, all eight compute warps ----
    bf16* h1 = hbuf;
    bf16* h2 = hbuf + GS * HP;
    if (!is_chain && !is_io) {
#pragma unroll
      for (int j = 0; j < CNT; ++j) {
        if (nt0 + j < nt1) {
          const int c = (nt0 + j) * 8 + 2 * t4;
          *reinterpret_cast<uint32_t*>(h1 + g * HP + c) =
              frag_pack(fmaxf(skip[j][0] + bias3[c], 0.f), fmaxf(skip[j][1] + bias3[c + 1], 0.f));
          *reinterpret_cast<uint32_t*>(h1 + (g + 8) * HP + c) =
              frag_pack(fmaxf(skip[j][2] + bias3[c], 0.f), fmaxf(skip[j][3] + bias3[c + 1], 0.f));
        }
      }
    }
    cbar();
    float acc[(NHP > 1 ? NHP : 1) * 4][4];
    auto dense = [&](const bf16* A, auto kc, auto nh) {
      constexpr int KC = decltype(kc)::value, NH = decltype(nh)::value;
#pragma unroll
      for (int i = 0; i < NH * 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int c8 = 0; c8 < KC; ++c8) {
        const unsigned char* w = slot_wait();
        uint32_t af[4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          lda_frag(af, A, HP, c8 * 32 + ks * 16, lane);
#pragma unroll
          for (int h = 0; h < NH; ++h) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint2 b = ldb_frag(w, h * 32 + warp * 4 + j, ks, 2, lane);
              mma16816(acc[h * 4 + j], af, b.x, b.y);
            }
          }
        }
        slot_release();
      }
    };
    dense(h1, std::integral_constant<int, S / 32>{}, std::integral_constant<int, NHP>{});
#pragma unroll
    for (int hj = 0; hj < NHP * 4; ++hj) {
      const int c = ((hj >> 2) * 32 + warp * 4 + (hj & 3)) * 8 + 2 * t4;
      *reinterpret_cast<uint32_t*>(h2 + g * HP + c) =
          frag_pack(fmaxf(acc[hj][0] + bias3[S + c], 0.f), fmaxf(acc[hj][1] + bias3[S + c + 1], 0.f));
      *reinterpret_cast<uint32_t*>(h2 + (g + 8) * HP + c) =
          frag_pack(fmaxf(acc[hj][2] + bias3[S + c], 0.f), fmaxf(acc[hj][3] + bias3[S + c + 1], 0.f));
    }
    cbar();
    dense(h2, std::integral_constant<int, P / 32>{}, std::integral_constant<int, 1>{});
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (warp * 4 + j) * 8 + 2 * t4;
      *reinterpret_cast<float2*>(lgbuf + g * LP + c) = make_float2(acc[j][0] + bias3[S + P + c], acc[j][1] + bias3[S + P + c + 1]);
      *reinterpret_cast<float2*>(lgbuf + (g + 8) * LP + c) = make_float2(acc[j][2] + bias3[S + P + c], acc[j][3] + bias3[S + P + c + 1]);
    }
    cbar();
    // ---- sampling (imodel.py:167-187) + teacher forcing (imodel.py:260-267): warp w -> streams 2w, 2w+1, in lock step ----
    {
      const int sa = warp * 2, sb = warp * 2 + 1;
      if (a.logits_out != nullptr) {
#pragma unroll
        for (int k = 0; k < 2; ++k)
          if (s0 + sa + k < a.n_streams)
            for (int q = lane; q < Q; q += 32)
              a.logits_out[((int64_t)(s0 + sa + k) * a.n_steps + step) * Q + q] = lgbuf[(sa + k) * LP + q];
      }
      int ra, rb;
      warp_sample2(lgbuf + sa * LP, lgbuf + sb * LP, ua, ub, ra, rb);
      if (lane < 2) {
        const int s = sa + lane, samp = lane == 0 ? ra : rb;
        if (s0 + s < a.n_streams) {
          a.out[(int64_t)(s0 + s) * a.n_steps + step] = samp;
          code_s[s] = (t < a.n_teacher) ? a.teacher[t] : samp;
        }
      }
    }
    cbar();
    G3_MARK(9)
  }
#undef G3_MARK
  if (prof && lane == 0)
    for (int i = 0; i < 10; ++i) a.trace[i] = pacc[i];
  // persist the state a later launch continues from: the pending codes (the rings already hold every x[t-dil])
  if (tid < GS && s0 + tid < a.n_streams) a.codes[s0 + tid] = code_s[tid];
}

// ---- host ------------------------------------------------------------------------------------------------
bool gen2_supported(const wn_model* m) {
  static const bool disabled = getenv("WN_DISABLE_GEN2") != nullptr;
  const wn_arch& a = m->a;
  return !disabled && a.n_res == 32 && a.n_dil == 32 && (a.n_skip == 256 || a.n_skip == 512) &&
         (a.n_post == 256 || a.n_post == 512) && a.n_quant == 256 && m->L >= 2 && m->L <= 64;
}
int64_t gen2_blob_bytes(const wn_model* m) { return gen2_layout(m->L, m->a.n_skip, m->a.n_post).total; }

int gen2_prepare(wn_model* m, const float* d_params, unsigned char* blob, cudaStream_t st) {
  const int S = m->a.n_skip, P = m->a.n_post;
  const Gen2Layout g = gen2_layout(m->L, S, P);
  k_gen2_prep<<<m->L + S / 32 + P / 32 + 1, 256, 0, st>>>(d_params, m->d_layers, m->L, m->off_pre, m->off_pre_b,
                                                           m->off_post1, m->off_post1_b, m->off_post2, m->off_post2_b,
                                                           blob, g, S, P);
  WN_LAUNCH_CHECK();
  return WN_OK;
}

template <int S, int P, bool GC>
static int gen2_launch(const Gen2Args& a, int n_streams, cudaStream_t st) {
  using namespace g2;
  // generation 2 by default; WN_GEN3=1 selects the generation-3 kernel (the chain on one warp): measured 37 us/step
  // against 25.4 us for generation 2 at 256 streams (its publish / wait pairs cost more than the barriers they replace)
  static const bool use_gen2 = getenv("WN_GEN3") == nullptr;
  constexpr size_t smem = (size_t)n_slots(S, P) * slot_bytes(S, P) + 257 * 32 * 2 +
                          (size_t)(4 * GS * XP + 2 * GS * hp(S, P)) * 2 + (size_t)GS * LP * 4 +
                          (size_t)OLD_W * GS * XP * 2 + (size_t)(S + P + Q) * 4 + 1024;
  static_assert(smem <= 226 * 1024, "generator shared memory");
  auto kern = use_gen2 ? k_gen2<S, P, GC> : k_gen3<S, P, GC>;
  WN_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps(PROF_GEN, st);
  kern<<<(n_streams + GS - 1) / GS, THREADS, smem, st>>>(a);
  WN_LAUNCH_CHECK();
  return WN_OK;
}

int gen2_run(wn_model* m, const unsigned char* blob, const int64_t* ring_off, bf16* rings, int32_t* codes,
             int n_streams, int64_t t0, int n_steps, uint64_t seed, const int32_t* teacher, int n_teacher, int32_t* out,
             float* logits, const float* gcproj, cudaStream_t st) {
  const int S = m->a.n_skip, P = m->a.n_post;
  Gen2Args a;
  memset(&a, 0, sizeof(a));
  a.blob = blob; a.g = gen2_layout(m->L, S, P); a.layers = m->d_layers; a.ring_off = ring_off; a.rings = rings;
  a.codes = codes; a.teacher = teacher; a.n_teacher = teacher ? n_teacher : 0; a.out = out; a.logits_out = logits;
  a.trace = g_trace_buf;
  a.t0 = t0; a.seed = seed; a.n_streams = n_streams; a.n_steps = n_steps; a.L = m->L;
  a.gcproj = gcproj;
  const bool gc = gcproj != nullptr;
#define WN_GEN2(S_, P_)                                                              \
  if (S == S_ && P == P_)                                                            \
    return gc ? gen2_launch<S_, P_, true>(a, n_streams, st) : gen2_launch<S_, P_, false>(a, n_streams, st);
  WN_GEN2(256, 256)
  WN_GEN2(512, 512)
  WN_GEN2(256, 512)
  WN_GEN2(512, 256)
#undef WN_GEN2
  set_error("gen2_run: unsupported n_skip / n_post %d / %d", S, P);
  return WN_ERR_UNSUPPORTED;
}

}  // namespace wn
