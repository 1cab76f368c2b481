// Training hot path, tcgen05 generation: TMA-staged operands, tcgen05.mma with TMEM accumulators,
// fused epilogues.  One CTA per 128-timestep tile; warp-specialised roles:
//   warp 0     TMA producer (one elected lane)
//   warp 1     MMA issuer  (one elected lane)
//   warps 2..5 epilogue: thread <-> TMEM lane <-> timestep row of the tile
//
// k_post_fwd_umma  (reference tmodel.py:321-324 skip sum, :187-215 post-net, :228-249 loss):
//   acc0[128 x S]  = Z[128 x L*D] . SKIPcat[L*D x S]        (concat-K over the layers == sum_l z_l . Ws_l)
//   h1 = bf16(relu(acc0 + sum_l bs_l))                     -> smem (A operand of the next MMA) + TMA store
//   acc1[128 x P]  = h1 . POST1 ;  h2 = bf16(relu(acc1 + b1)) -> smem + TMA store
//   acc0[128 x Q]  = h2 . POST2 ;  logits = acc0 + b2 (never leave the chip unless asked)
//   masked softmax cross entropy per row; dlogits = (softmax - onehot) * mask -> smem -> TMA store
// B operands are the transposed weights (N x K, K-major) prepared once per step by k_prep_umma_weights.
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "umma.cuh"

namespace wn {

using namespace umma;

constexpr int UM = 128;            // rows (timesteps) per tile
constexpr int UKB = 64;            // K elements per pipeline block (one 128-byte swizzle span of bf16)
constexpr int USTAGES = 3;
constexpr int UA_BYTES = UM * 128;   // 16 KB
constexpr int UB_BYTES = 256 * 128;  // 32 KB (N <= 256)
constexpr int UH_BYTES = 4 * UA_BYTES;  // activation tile [128 x 256] bf16 as four K blocks
constexpr int UPOST_THREADS = 192;

struct PostUmmaArgs {
  const float* params;
  const float* skip_bias;
  int64_t off_post1_b, off_post2_b;
  const int32_t* wav;
  const int32_t* ids;
  double* stats;
  float* logits_out;
  int T, S, P, Q, LD, use_bias;
  int64_t rows;
  uint32_t *hm1, *hm2;  // relu masks of h1 / h2 as bits, [rows padded to 128][8] (model.h)
  long long* trace;  // wn_debug_trace(buf, -2): in-kernel timeline of CTA 0 (tools/trace_layer.py postfwd)
};

// ---- weight preparation -------------------------------------------------------------------------
// wsT[s][l*D+d] = SKIP_l[d][s] ; wsCat[l*D+d][s] = SKIP_l[d][s] ; w1T[p][s] = POST1[s][p] ; w2T[q][p] = POST2[p][q]
__global__ void k_prep_umma_weights(const float* __restrict__ p, const LayerDesc* __restrict__ layers, int L, int D,
                                    int S, int P, int Q, int64_t off_post1, int64_t off_post2,
                                    bf16* __restrict__ wsT, bf16* __restrict__ wsCat, bf16* __restrict__ w1T,
                                    bf16* __restrict__ w2T) {
  const int64_t LD = (int64_t)L * D;
  const int64_t n_ws = LD * S, n_w1 = (int64_t)S * P, n_w2 = (int64_t)P * Q;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_ws + n_w1 + n_w2;
       i += (int64_t)gridDim.x * blockDim.x) {
    if (i < n_ws) {
      const int64_t k = i / S;
      const int s = (int)(i % S);
      const int l = (int)(k / D), d = (int)(k % D);
      const bf16 v = f2bf(p[layers[l].skip + (int64_t)d * S + s]);
      wsCat[i] = v;
      wsT[(int64_t)s * LD + k] = v;
    } else if (i < n_ws + n_w1) {
      const int64_t j = i - n_ws;
      const int s = (int)(j / P), pp = (int)(j % P);
      w1T[(int64_t)pp * S + s] = f2bf(p[off_post1 + j]);
    } else {
      const int64_t j = i - n_ws - n_w1;
      const int pp = (int)(j / Q), q = (int)(j % Q);
      w2T[(int64_t)q * P + pp] = f2bf(p[off_post2 + j]);
    }
  }
}

// ---- helpers ------------------------------------------------------------------------------------
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// write 32 consecutive columns [c0, c0+32) of `row` (already converted to packed bf16) into the
// K-major SW128 activation tile: K block = c0/64, 64 bytes = four 16-byte chunks
__device__ __forceinline__ void htile_store32(unsigned char* htile, int row, int c0, const uint32_t (&pk)[16]) {
  unsigned char* blk = htile + (c0 >> 6) * UA_BYTES;
  const int byte0 = (c0 & 63) * 2;
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    const uint32_t off = swizzled_offset((uint32_t)row, (uint32_t)(byte0 + ch * 16), 128);
    *reinterpret_cast<uint4*>(blk + off) = make_uint4(pk[ch * 4 + 0], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
  }
}

// relu of 32 pre-activations (accumulator + bias), packed to bf16, and the mask word: bit j = 1 unless x_j carries a sign
// bit, i.e. "h_j > 0" (x = +0 counts as positive: measure zero, and where whole columns are exactly zero -- channel
// padding -- the weights that would carry a gradient are zero too).  Two instructions per column: the sign bits are
// shifted into the word one after the other.
__device__ __forceinline__ uint32_t relu_pack32(const uint32_t (&v)[32], const float* bias, uint32_t (&pk)[16]) {
  uint32_t nb = 0u;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float x0 = __uint_as_float(v[2 * j]) + bias[2 * j], x1 = __uint_as_float(v[2 * j + 1]) + bias[2 * j + 1];
    nb = (nb >> 1) | (__float_as_uint(x0) & 0x80000000u);
    nb = (nb >> 1) | (__float_as_uint(x1) & 0x80000000u);
    pk[j] = pack_bf16x2(fmaxf(x0, 0.f), fmaxf(x1, 0.f));
  }
  return ~nb;
}

constexpr int UPOST_P_THREADS = 320;  // producer warp, MMA warp, 8 epilogue warps

__device__ __forceinline__ void epi_bar_sync256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Persistent: one CTA per SM loops over 128-row tiles.  TMEM holds two 256-column accumulators whose roles rotate
// with the tile parity p: GEMM1 -> buf[p], GEMM2 -> buf[p^1], GEMM3 -> buf[p]; therefore the MMA warp can issue
// tile i+1's skip GEMM (the long one, K = L*D) while the epilogue warps are still busy with tile i's loss.
// PAIR: the same kernel on CTA pairs (clusters of two, tcgen05 cta_group::2).  The kernel runs at the rate the L2 delivers
// its operands (~35 B/clk/SM) and three quarters of them are weights re-streamed for every 128-row tile: in a pair each
// CTA loads only HALF of every weight block (its N / 2 rows); the leader's M = 256 MMAs read both CTAs' shared memory
// and fill both CTAs' tensor memory; everything else (A tiles, accumulators, epilogues, stores) stays per CTA.  What
// the MMA issuer waits for -- stage landed, accumulator drained, activation tile written -- it waits for from BOTH
// CTAs (TMA completions and arrivals directed at the leader's barriers); what it signals, it multicasts.
template <bool PAIR>
__global__ void __launch_bounds__(UPOST_P_THREADS, 1)
k_post_fwd_umma(const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_wsT,
                const __grid_constant__ CUtensorMap map_w1T, const __grid_constant__ CUtensorMap map_w2T,
                const __grid_constant__ CUtensorMap map_h1, const __grid_constant__ CUtensorMap map_h2,
                const __grid_constant__ CUtensorMap map_dlog, PostUmmaArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int NSTG = PAIR ? 4 : USTAGES;                         // ring depth
  constexpr int B_BYTES = PAIR ? UB_BYTES / 2 : UB_BYTES;           // this CTA's part of a weight block
  unsigned char* stage_a = smem;                                  // NSTG x 16 KB
  unsigned char* stage_b = smem + NSTG * UA_BYTES;                // NSTG x 32 (16) KB
  unsigned char* htile = stage_b + NSTG * B_BYTES;                // 64 KB
  uint32_t* mtile = reinterpret_cast<uint32_t*>(htile + UH_BYTES);  // 4 KB: [128 rows][8 words] relu mask bits of the tile
  __shared__ __align__(8) uint64_t full_bar[NSTG], empty_bar[NSTG], acc_full[2], acc_empty[2], h_ready[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[3 * 256];  // skip-bias sum | POST1_BIAS | POST2_BIAS (broadcast reads in the epilogues)
  __shared__ float x_mx[2][128], x_sum[2][128], x_vl[2][128];  // softmax partials exchanged between the two
  __shared__ int x_arg[2][128];                                 // column halves of a row

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.S, P = a.P, Q = a.Q;
  const int n_tiles = (int)((a.rows + UM - 1) / UM);
  // work unit: a tile (one CTA) or a pair of consecutive tiles (one cluster; CTA `rank` takes tile 2 * pair + rank, which
  // may lie beyond the batch: TMA zero fill in, clipped stores out)
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, n_units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int n_work = PAIR ? (n_tiles + 1) / 2 : n_tiles;
  const int n_my = (n_work - unit0 + n_units - 1) / n_units;
  auto tile_row0 = [&](int i) -> int64_t {
    const int64_t u = (int64_t)unit0 + (int64_t)i * n_units;
    return (PAIR ? 2 * u + rank : u) * UM;
  };
  for (int i = tid; i < 3 * 256; i += UPOST_P_THREADS) {
    float v = 0.f;
    if (a.use_bias) {
      const int which = i >> 8, c = i & 255;
      if (which == 0 && c < S) v = a.skip_bias[c];
      if (which == 1 && c < P) v = a.params[a.off_post1_b + c];
      if (which == 2 && c < Q) v = a.params[a.off_post2_b + c];
    }
    bias_s[i] = v;
  }
  const int nkb1 = (a.LD + UKB - 1) / UKB, nkb2 = S / UKB, nkb3 = P / UKB;
  PostTracer tr;
  tr.init(a.trace, warp, blockIdx.x == 0 && lane == 0);

  if (tid == 0) {
    for (int i = 0; i < NSTG; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], PAIR ? 2 : 256);  // pair: one arrival per CTA (its elected epilogue thread), at the leader
      mbar_init(&h_ready[i], PAIR ? 2 : 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&map_z);
    tma_prefetch_desc(&map_wsT);
    tma_prefetch_desc(&map_w1T);
    tma_prefetch_desc(&map_w2T);
  }
  if (warp == 1) {
    if constexpr (PAIR) tmem_alloc_pair(&tmem_base_s, 512); else tmem_alloc(&tmem_base_s, 512);
  }
  tc_fence_before_sync();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // both CTAs' barriers exist before a remote completion or arrival can hit them
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      auto load = [&](void* dst, const CUtensorMap* mp, uint64_t* bar, int c0, int c1) {
        if constexpr (PAIR) tma_load_2d_pair(dst, mp, bar, c0, c1); else tma_load_2d(dst, mp, bar, c0, c1);
      };
      for (int i = 0; i < n_my; ++i) {
        const int row0 = (int)tile_row0(i);
        for (int kb = 0; kb < nkb1 + nkb2 + nkb3; ++kb, ++it) {
          const int st = it % NSTG;
          tr.ev(1, it);
          mbar_wait(&empty_bar[st], ((uint32_t)(it / NSTG) & 1u) ^ 1u);
          tr.ev(2, it);
          unsigned char* sa = stage_a + st * UA_BYTES;
          unsigned char* sb = stage_b + st * B_BYTES;
          // bytes of the WHOLE stage (pair: both CTAs' loads complete on the leader's barrier, which alone expects them)
          const int nrows = kb < nkb1 ? S : kb < nkb1 + nkb2 ? P : Q;  // N of this contraction
          const uint32_t bytes = (uint32_t)(nrows * 128 + (kb < nkb1 ? (PAIR ? 2 : 1) * UA_BYTES : 0));
          if (!PAIR || rank == 0) mbar_expect_tx(&full_bar[st], bytes);
          const int brow = PAIR ? rank * (nrows / 2) : 0;  // this CTA's rows of the weight block
          if (kb < nkb1) {
            load(sa, &map_z, &full_bar[st], kb * UKB, row0);
            load(sb, &map_wsT, &full_bar[st], kb * UKB, brow);
          } else if (kb < nkb1 + nkb2) {
            load(sb, &map_w1T, &full_bar[st], (kb - nkb1) * UKB, brow);
          } else {
            load(sb, &map_w2T, &full_bar[st], (kb - nkb1 - nkb2) * UKB, brow);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (pair: the leader's only) =====
    if (lane == 0 && rank == 0) {
      int it = 0;
      uint32_t use[2] = {0, 0};  // number of contractions issued into each accumulator buffer so far
      constexpr int MM = PAIR ? 2 * UM : UM;
      const uint32_t idesc1 = make_idesc_bf16(MM, S), idesc2 = make_idesc_bf16(MM, P), idesc3 = make_idesc_bf16(MM, Q);
      auto commit = [&](uint64_t* bar) {
        if constexpr (PAIR) mma_commit_pair(bar, 3); else mma_commit(bar);
      };
      auto gemm = [&](int buf, int nkb, uint32_t idesc, bool a_from_htile) {
        tr.ev(3, it);
        mbar_wait(&acc_empty[buf], (use[buf] & 1u) ^ 1u);  // the previous contraction of this buffer has been drained
        tr.ev(4, it);
        tc_fence_after_sync();
        const uint32_t acc = tmem_base + (uint32_t)buf * 256;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int st = it % NSTG;
          mbar_wait(&full_bar[st], (uint32_t)(it / NSTG) & 1u);
          tr.ev(16, it);
          tc_fence_after_sync();
          const uint32_t sa = a_from_htile ? smem_u32(htile + kb * UA_BYTES) : smem_u32(stage_a + st * UA_BYTES);
          const uint32_t sb = smem_u32(stage_b + st * B_BYTES);
#pragma unroll
          for (int k = 0; k < UKB / 16; ++k) {
            const uint64_t ad = make_kmajor_desc(sa, 128, k * 32), bd = make_kmajor_desc(sb, 128, k * 32);
            if constexpr (PAIR) mma_bf16_ss_pair(acc, ad, bd, idesc, (kb | k) != 0);
            else mma_bf16_ss(acc, ad, bd, idesc, (kb | k) != 0);
          }
          commit(&empty_bar[st]);
        }
        commit(&acc_full[buf]);
        tr.ev(17, it);
        ++use[buf];
      };
      for (int i = 0; i < n_my; ++i) {
        const int p = i & 1;
        gemm(p, nkb1, idesc1, false);
        mbar_wait(&h_ready[0], (uint32_t)i & 1u);
        tc_fence_after_sync();
        gemm(p ^ 1, nkb2, idesc2, true);
        mbar_wait(&h_ready[1], (uint32_t)i & 1u);
        tc_fence_after_sync();
        gemm(p, nkb3, idesc3, true);
      }
    }
  } else {
    // ===== epilogue warps: thread <-> (row, column half) =====
    const int e = warp - 2, q4 = warp & 3, half = e >> 2;
    const int r = q4 * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const bool elected = (warp == 2 && lane == 0);
    uint32_t use[2] = {0, 0};
    uint32_t v[32];
    uint32_t pk[16];
    float acc_x = 0.f, acc_n = 0.f, acc_d = 0.f;  // loss statistics over all tiles of this CTA
    int n_epi = 0;  // epilogue phases so far (trace only)
    auto acc_wait = [&](int buf) -> uint32_t {
      tr.ev(5, n_epi);
      mbar_wait(&acc_full[buf], use[buf] & 1u);
      tr.ev(6, n_epi);
      tc_fence_after_sync();
      return tmem_base + (uint32_t)buf * 256 + lane_sel;
    };
    auto acc_release = [&](int buf) {
      tc_fence_before_sync();
      if constexpr (!PAIR) mbar_arrive(&acc_empty[buf]);  // (pair: one arrival per CTA, by the elected thread after the barrier)
      tr.ev(7, n_epi++);
      ++use[buf];
    };
    for (int i = 0; i < n_my; ++i) {
      const int p = i & 1;
      const int64_t row0 = tile_row0(i);
      const int64_t row = row0 + r;
      // ---- h1 = relu(skip_sum + bias) ----
      {
        const uint32_t acc = acc_wait(p);
        if (elected) tma_store_wait_read<0>();  // the previous tile's dlogits store has finished reading the tile
        epi_bar_sync256();
        const int cb = half * (S / 2);
        for (int c0 = cb; c0 < cb + S / 2; c0 += 32) {
          tmem_ld_32x32b_x32(acc + (uint32_t)c0, v);
          tmem_ld_wait();
          mtile[r * 8 + (c0 >> 5)] = relu_pack32(v, bias_s + c0, pk);
          htile_store32(htile, r, c0, pk);
        }
        acc_release(p);
        fence_proxy_async_smem();
        epi_bar_sync256();
        if (elected) {
          if constexpr (PAIR) {
            mbar_arrive_remote(&acc_empty[p], 0);
            mbar_arrive_remote(&h_ready[0], 0);
          } else {
            mbar_arrive(&h_ready[0]);
          }
          for (int kb = 0; kb < nkb2; ++kb) tma_store_2d(&map_h1, htile + kb * UA_BYTES, kb * UKB, (int)row0);
          if (row0 < a.rows) bulk_store_1d(a.hm1 + (size_t)row0 * 8, mtile, 128 * 32);  // (the mask buffers are padded to whole tiles)
          tma_store_commit();
        }
      }
      // ---- h2 = relu(h1 . POST1 + b1) ----
      {
        const uint32_t acc = acc_wait(p ^ 1);   // also: the contraction reading h1 from the tile has completed
        if (elected) tma_store_wait_read<0>();  // the h1 store has finished reading the tile
        epi_bar_sync256();
        const int cb = half * (P / 2);
        for (int c0 = cb; c0 < cb + P / 2; c0 += 32) {
          tmem_ld_32x32b_x32(acc + (uint32_t)c0, v);
          tmem_ld_wait();
          mtile[r * 8 + (c0 >> 5)] = relu_pack32(v, bias_s + 256 + c0, pk);
          htile_store32(htile, r, c0, pk);
        }
        acc_release(p ^ 1);
        fence_proxy_async_smem();
        epi_bar_sync256();
        if (elected) {
          if constexpr (PAIR) {
            mbar_arrive_remote(&acc_empty[p ^ 1], 0);
            mbar_arrive_remote(&h_ready[1], 0);
          } else {
            mbar_arrive(&h_ready[1]);
          }
          for (int kb = 0; kb < nkb3; ++kb) tma_store_2d(&map_h2, htile + kb * UA_BYTES, kb * UKB, (int)row0);
          if (row0 < a.rows) bulk_store_1d(a.hm2 + (size_t)row0 * 8, mtile, 128 * 32);
          tma_store_commit();
        }
      }
      // ---- logits, masked softmax cross entropy, dlogits ----
      {
        const uint32_t acc = acc_wait(p);
        const bool in_range = row < a.rows;
        const int b = in_range ? (int)(row / a.T) : 0, t = in_range ? (int)(row % a.T) : 0;
        const bool has_next = in_range && (t + 1 < a.T);
        const bool valid = has_next && (a.ids[(size_t)b * a.T + t + 1] != 0);  // tmodel.py:232
        // tmodel.py:230 + :64: tf.one_hot of an out-of-range code is an all-zero row, so the cross entropy of that
        // position is 0, dlogits = softmax - 0 (TF's fused xent kernel returns softmax - labels) and argmax(label) = 0
        const int label_raw = has_next ? a.wav[(size_t)b * a.T + t + 1] : 0;
        const bool label_ok = label_raw >= 0 && label_raw < Q;
        const int label = label_ok ? label_raw : -1;
        const int cb = half * (Q / 2);
        // pass 1 (online softmax over this thread's column half): running max, sum of exp, argmax, label logit
        float mx = -INFINITY, sum = 0.f, vl = 0.f;
        int arg = 0;
        for (int c0 = cb; c0 < cb + Q / 2; c0 += 32) {
          tmem_ld_32x32b_x32(acc + (uint32_t)c0, v);
          tmem_ld_wait();
          float cm = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = __uint_as_float(v[j]) + bias_s[512 + c0 + j];
            v[j] = __float_as_uint(x);
            if (x > mx && x > cm) arg = c0 + j;  // first maximum: strictly greater than everything before
            cm = fmaxf(cm, x);
            if (c0 + j == label) vl = x;
            if (a.logits_out != nullptr && in_range) a.logits_out[(size_t)row * Q + c0 + j] = x;
          }
          const float nm = fmaxf(mx, cm);
          float cs = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) cs += __expf(__uint_as_float(v[j]) - nm);
          sum = sum * __expf(mx - nm) + cs;
          mx = nm;
        }
        // exchange the partials between the two column halves of the row
        x_mx[half][r] = mx;
        x_sum[half][r] = sum;
        x_arg[half][r] = arg;
        x_vl[half][r] = vl;
        if (elected) tma_store_wait_read<0>();  // the h2 store has finished reading the tile
        epi_bar_sync256();
        const float om = x_mx[half ^ 1][r], os = x_sum[half ^ 1][r];
        const float M = fmaxf(mx, om);
        const float tot = sum * __expf(mx - M) + os * __expf(om - M);
        const float inv = 1.f / tot;
        for (int c0 = cb; c0 < cb + Q / 2; c0 += 32) {  // pass 2: dlogits = (softmax - onehot) * mask
          tmem_ld_32x32b_x32(acc + (uint32_t)c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float x0 = __uint_as_float(v[2 * j]) + bias_s[512 + c0 + 2 * j];
            const float x1 = __uint_as_float(v[2 * j + 1]) + bias_s[512 + c0 + 2 * j + 1];
            const float g0 = __expf(x0 - M) * inv - ((c0 + 2 * j) == label ? 1.f : 0.f);
            const float g1 = __expf(x1 - M) * inv - ((c0 + 2 * j + 1) == label ? 1.f : 0.f);
            pk[j] = valid ? pack_bf16x2(g0, g1) : 0u;
          }
          htile_store32(htile, r, c0, pk);
        }
        acc_release(p);
        fence_proxy_async_smem();
        epi_bar_sync256();
        if (elected) {
          if constexpr (PAIR) mbar_arrive_remote(&acc_empty[p], 0);
          for (int kb = 0; kb < Q / UKB; ++kb) tma_store_2d(&map_dlog, htile + kb * UA_BYTES, kb * UKB, (int)row0);
          tma_store_commit();
        }
        if (half == 0 && valid) {  // one thread per row owns the statistics
          // argmax over the whole row: smallest index among equal maxima (tf.argmax)
          const int garg = (x_mx[1][r] > x_mx[0][r]) ? x_arg[1][r] : x_arg[0][r];
          const float gvl = label < Q / 2 ? x_vl[0][r] : x_vl[1][r];
          acc_x += label_ok ? __logf(tot) + M - gvl : 0.f;
          acc_n += 1.f;
          acc_d += fabsf((float)((label_ok ? label : 0) - garg));
        }
      }
    }
    acc_x = warp_sum(acc_x);
    acc_n = warp_sum(acc_n);
    acc_d = warp_sum(acc_d);
    if (lane == 0 && acc_n != 0.f) {
      atomicAdd(a.stats + WN_STAT_XENT_SUM, (double)acc_x);
      atomicAdd(a.stats + WN_STAT_N_VALID, (double)acc_n);
      atomicAdd(a.stats + WN_STAT_DIFF_SUM, (double)acc_d);
    }
    if (elected) tma_store_wait_all<0>();
  }
  tc_fence_before_sync();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // neither CTA leaves while the other may still be served by its memories
  if (warp == 1) {
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// =====================================================================================================
// k_post_bwd_umma: dlogits -> dp1 = (dlogits . POST2^T) * (h2 > 0) -> dskip = (dp1 . POST1^T) * (h1 > 0)
//                  -> dz[:, l*D:(l+1)*D] = dskip . SKIP_l^T for every layer (N = L*D in chunks of 256),
// plus the three bias gradients (column sums of dlogits / dp1 / dskip) computed from the shared-memory
// tiles by the epilogue warps while the tensor core works on the next contraction.
// B operands are the weights in their natural [N][K] row-major layout (POST2 [P][Q], POST1 [S][P]) and the
// layer-stacked SKIP matrix wsCat [L*D][S].
// =====================================================================================================
constexpr int UBSTAGES = 3;

struct PostBwdUmmaArgs {
  const bf16* h1;
  const bf16* h2;
  float* g_post2_b;   // or nullptr
  float* g_post1_b;
  float* g_skip_b0;   // SKIP_BIAS of layer 0 (broadcast to the other layers afterwards)
  int S, P, Q, LD, D;
  int64_t rows;
  const uint32_t *hm1, *hm2;  // relu masks of h1 / h2 as bits (written by k_post_fwd_umma)
  long long* trace;  // wn_debug_trace(buf, -2) (tools/trace_layer.py postbwd)
};

// partial column sums of a [128 x ncols] K-major SW128 tile: thread et of the 256 epilogue threads owns the column
// pair et / 2 and the row half et % 2; accumulated in registers across all tiles of the CTA
__device__ __forceinline__ void tile_colsum_acc(const unsigned char* tile, int ncols, int et, int rows_valid, float (&acc)[2]) {
  const int j = et >> 1;
  if (2 * j >= ncols) return;
  const unsigned char* blk = tile + ((2 * j) >> 6) * UA_BYTES;
  const uint32_t byte_in_row = (uint32_t)((2 * j) & 63) * 2;
  const int rb = (et & 1) * 64, re = min(rows_valid, rb + 64);
  float s0 = 0.f, s1 = 0.f;
  for (int r = rb; r < re; ++r) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(blk + swizzled_offset((uint32_t)r, byte_in_row, 128));
    s0 += __uint_as_float(w << 16);
    s1 += __uint_as_float(w & 0xffff0000u);
  }
  acc[0] += s0;
  acc[1] += s1;
}

// relu mask from 32 stored activations (64 bytes of one row), applied while packing to bf16
// the same with the mask as 32 bits (bit j <-> column j), written by the fused forward: one word instead of 64 bytes
__device__ __forceinline__ void mask_pack32_bits(const uint32_t (&v)[32], uint32_t bits, uint32_t (&pk)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float m0 = (bits >> (2 * j)) & 1u ? 1.f : 0.f, m1 = (bits >> (2 * j + 1)) & 1u ? 1.f : 0.f;
    pk[j] = pack_bf16x2(__uint_as_float(v[2 * j]) * m0, __uint_as_float(v[2 * j + 1]) * m1);
  }
}
__device__ __forceinline__ uint32_t ldg_u32_pred(const uint32_t* p, bool pred) {
  uint32_t v;
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.u32 q, %2, 0;\n"
      "mov.u32 %0, 0;\n"
      "@q ld.global.nc.u32 %0, [%1];\n"
      "}\n"
      : "=&r"(v)
      : "l"(p), "r"((uint32_t)pred));
  return v;
}
__device__ __forceinline__ void mask_pack32(const uint32_t (&v)[32], const bf16* hrow, uint32_t (&pk)[16]) {
  const uint4* h4 = reinterpret_cast<const uint4*>(hrow);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 hv = __ldg(h4 + q);
    const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = q * 4 + e;  // pair index: columns 2j, 2j+1
      const float m0 = (hw[e] & 0x0000ffffu) != 0u && !(hw[e] & 0x00008000u) ? 1.f : 0.f;  // h > 0 (h >= 0 always: relu output)
      const float m1 = (hw[e] & 0xffff0000u) != 0u && !(hw[e] & 0x80000000u) ? 1.f : 0.f;
      pk[j] = pack_bf16x2(__uint_as_float(v[2 * j]) * m0, __uint_as_float(v[2 * j + 1]) * m1);
    }
  }
}

// Persistent: one CTA per SM loops over 128-row tiles; warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue
// (thread <-> (row, column half)).  The weight stream (B operands) runs through a 3-stage ring that never drains
// between tiles; the next tile's dlogits are fetched as soon as the last dz-chunk contraction has released tile0, so
// its first contraction overlaps the tail epilogues of the current tile.
__global__ void __launch_bounds__(UPOST_P_THREADS, 1)
k_post_bwd_umma(const __grid_constant__ CUtensorMap map_dlog, const __grid_constant__ CUtensorMap map_w2,
                const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_wsCat,
                const __grid_constant__ CUtensorMap map_dp1, const __grid_constant__ CUtensorMap map_dskip,
                const __grid_constant__ CUtensorMap map_dz, PostBwdUmmaArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* tile0 = smem;                          // 64 KB: dlogits, later dskip
  unsigned char* tile1 = smem + UH_BYTES;               // 64 KB: dp1, later dz staging
  unsigned char* stage_b = smem + 2 * UH_BYTES;         // UBSTAGES x 32 KB
  __shared__ __align__(8) uint64_t full_bar[UBSTAGES], empty_bar[UBSTAGES], a_full, t0_free, acc_full[2], acc_empty[2],
      t_ready[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.S, P = a.P, Q = a.Q, LD = a.LD;
  const int n_tiles = (int)((a.rows + UM - 1) / UM);
  const int n_my = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int nkb4 = Q / UKB, nkb5 = P / UKB, nkb6 = S / UKB;
  const int nchunk = (LD + 255) / 256;
  PostTracer tr;
  tr.init(a.trace, warp, blockIdx.x == 0 && lane == 0);

  if (tid == 0) {
    for (int i = 0; i < UBSTAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&a_full, 1);
    mbar_init(&t0_free, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 256);
      mbar_init(&t_ready[i], 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&map_dlog);
    tma_prefetch_desc(&map_w2);
    tma_prefetch_desc(&map_w1);
    tma_prefetch_desc(&map_wsCat);
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      auto load_b = [&](const CUtensorMap* mp, int k0, int n0, uint32_t bytes) {
        const int st = it % UBSTAGES;
        tr.ev(1, it);
        mbar_wait(&empty_bar[st], ((uint32_t)(it / UBSTAGES) & 1u) ^ 1u);
        tr.ev(2, it);
        mbar_expect_tx(&full_bar[st], bytes);
        tma_load_2d(stage_b + st * UB_BYTES, mp, &full_bar[st], k0, n0);
        ++it;
      };
      for (int i = 0; i < n_my; ++i) {
        const int row0 = ((int)blockIdx.x + i * (int)gridDim.x) * UM;
        if (i > 0) mbar_wait(&t0_free, (uint32_t)(i - 1) & 1u);  // the previous tile's dz contractions have released tile0
        mbar_expect_tx(&a_full, (uint32_t)(nkb4 * UA_BYTES));
        for (int kb = 0; kb < nkb4; ++kb) tma_load_2d(tile0 + kb * UA_BYTES, &map_dlog, &a_full, kb * UKB, row0);
        for (int kb = 0; kb < nkb4; ++kb) load_b(&map_w2, kb * UKB, 0, (uint32_t)(P * 128));
        for (int kb = 0; kb < nkb5; ++kb) load_b(&map_w1, kb * UKB, 0, (uint32_t)(S * 128));
        for (int c = 0; c < nchunk; ++c)
          for (int kb = 0; kb < nkb6; ++kb) load_b(&map_wsCat, kb * UKB, c * 256, (uint32_t)UB_BYTES);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int it = 0, use = 0;
      auto gemm = [&](const unsigned char* atile, int nkb, uint32_t idesc) {
        const int buf = use & 1, k_use = use >> 1;
        tr.ev(3, use);
        mbar_wait(&acc_empty[buf], ((uint32_t)k_use & 1u) ^ 1u);  // accumulator drained by the epilogue
        tr.ev(4, use);
        tc_fence_after_sync();
        const uint32_t acc = tmem_base + (uint32_t)buf * 256;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int st = it % UBSTAGES;
          mbar_wait(&full_bar[st], (uint32_t)(it / UBSTAGES) & 1u);
          tr.ev(16, it);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(atile + kb * UA_BYTES), sb = smem_u32(stage_b + st * UB_BYTES);
#pragma unroll
          for (int k = 0; k < UKB / 16; ++k)
            mma_bf16_ss(acc, make_kmajor_desc(sa, 128, k * 32), make_kmajor_desc(sb, 128, k * 32), idesc, (kb | k) != 0);
          mma_commit(&empty_bar[st]);
        }
        mma_commit(&acc_full[buf]);
        tr.ev(17, use);
        ++use;
      };
      for (int i = 0; i < n_my; ++i) {
        mbar_wait(&a_full, (uint32_t)i & 1u);
        tc_fence_after_sync();
        gemm(tile0, nkb4, make_idesc_bf16(UM, P));          // dp1 pre-activation gradient
        mbar_wait(&t_ready[1], (uint32_t)i & 1u);            // dp1 tile written
        tc_fence_after_sync();
        gemm(tile1, nkb5, make_idesc_bf16(UM, S));          // dskip pre-mask
        mbar_wait(&t_ready[0], (uint32_t)i & 1u);            // dskip tile written
        tc_fence_after_sync();
        for (int c = 0; c < nchunk; ++c) gemm(tile0, nkb6, make_idesc_bf16(UM, 256));  // dz chunks
        mma_commit(&t0_free);
      }
    }
  } else {
    const int e = warp - 2, q4 = warp & 3, half = e >> 2;
    const int r = q4 * 32 + lane;
    const int et = e * 32 + lane;  // 0..255 index among the epilogue threads
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const bool elected = (warp == 2 && lane == 0);
    uint32_t v[32];
    uint32_t pk[16];
    int use = 0;
    float b2acc[2] = {0.f, 0.f}, b1acc[2] = {0.f, 0.f}, bsacc[2] = {0.f, 0.f};
    auto acc_wait = [&]() -> uint32_t {
      const int buf = use & 1, k_use = use >> 1;
      tr.ev(5, use);
      mbar_wait(&acc_full[buf], (uint32_t)k_use & 1u);
      tr.ev(6, use);
      tc_fence_after_sync();
      return tmem_base + (uint32_t)buf * 256 + lane_sel;
    };
    auto acc_release = [&]() {
      tc_fence_before_sync();
      mbar_arrive(&acc_empty[use & 1]);
      tr.ev(7, use);
      ++use;
    };
    for (int i = 0; i < n_my; ++i) {
      const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)i * gridDim.x) * UM;
      const int64_t row = row0 + r;
      const bool in_range = row < a.rows;
      const int rows_valid = (int)min((int64_t)UM, a.rows - row0);
      // this row's relu masks (one bit per column, written by the forward): requested here, a contraction and more before
      // they are used.  (Read as the 2 x 256 bytes of h2 / h1 themselves inside the epilogue steps, every step exposed a
      // global-memory round trip on the tile's critical chain: in-kernel timeline 6600 + 7700 cycles for the two masked
      // epilogues against ~1100 for an unmasked one of the same size.)
      uint32_t mw2[4], mw1[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        mw2[k] = ldg_u32_pred(a.hm2 + (size_t)(in_range ? row : 0) * 8 + half * (P / 64) + k, in_range && 32 * k < P / 2);
        mw1[k] = ldg_u32_pred(a.hm1 + (size_t)(in_range ? row : 0) * 8 + half * (S / 64) + k, in_range && 32 * k < S / 2);
      }
      // POST2_BIAS gradient from the dlogits tile while the first contraction runs
      mbar_wait(&a_full, (uint32_t)i & 1u);
      if (a.g_post2_b != nullptr) tile_colsum_acc(tile0, Q, et, rows_valid, b2acc);
      // ---- dp1 ----
      {
        const uint32_t acc = acc_wait();
        if (elected) tma_store_wait_read<0>();  // the previous tile's last dz store has finished reading tile1
        epi_bar_sync256();
        const int cb = half * (P / 2);
        for (int c0 = cb; c0 < cb + P / 2; c0 += 32) {
          tmem_ld_32x32b_x32(acc + (uint32_t)c0, v);
          tmem_ld_wait();
          mask_pack32_bits(v, mw2[0], pk);  // (rows beyond the batch: bits 0 on an accumulator row of zeros)
          mw2[0] = mw2[1]; mw2[1] = mw2[2]; mw2[2] = mw2[3];
          htile_store32(tile1, r, c0, pk);
        }
        acc_release();
        fence_proxy_async_smem();
        epi_bar_sync256();
        if (elected) {
          for (int kb = 0; kb < nkb5; ++kb) tma_store_2d(&map_dp1, tile1 + kb * UA_BYTES, kb * UKB, (int)row0);
          tma_store_commit();
          mbar_arrive(&t_ready[1]);
        }
        if (a.g_post1_b != nullptr) tile_colsum_acc(tile1, P, et, rows_valid, b1acc);
      }
      // ---- dskip ----
      {
        const uint32_t acc = acc_wait();  // also: the first contraction (reading tile0) has completed
        const int cb = half * (S / 2);
        for (int c0 = cb; c0 < cb + S / 2; c0 += 32) {
          tmem_ld_32x32b_x32(acc + (uint32_t)c0, v);
          tmem_ld_wait();
          mask_pack32_bits(v, mw1[0], pk);
          mw1[0] = mw1[1]; mw1[1] = mw1[2]; mw1[2] = mw1[3];
          htile_store32(tile0, r, c0, pk);
        }
        acc_release();
        fence_proxy_async_smem();
        epi_bar_sync256();
        if (elected) {
          for (int kb = 0; kb < nkb6; ++kb) tma_store_2d(&map_dskip, tile0 + kb * UA_BYTES, kb * UKB, (int)row0);
          tma_store_commit();
          mbar_arrive(&t_ready[0]);
        }
        if (a.g_skip_b0 != nullptr) tile_colsum_acc(tile0, S, et, rows_valid, bsacc);
      }
      // ---- dz chunks: staging through tile1 (dp1 is dead: its contraction and its TMA store are complete) ----
      // dz is stored as per-layer planes [L][B*T][D] (dense rows for the layer backward): the staging tile is a row of
      // per-layer panels [128][D] whose rows are one swizzle span (2*D bytes)
      const int span = 2 * a.D, panel_bytes = UM * span;
      for (int c = 0; c < nchunk; ++c) {
        const uint32_t acc = acc_wait();
        if (elected) tma_store_wait_read<0>();  // previous stores (dp1 / dskip / previous chunk) done reading smem
        epi_bar_sync256();
        tr.ev(8, use);
        const int ncols = min(256, LD - c * 256);
        const int cb = half * 128, ce = min(ncols, cb + 128);
        for (int c0 = cb; c0 < ce; c0 += 32) {
          tmem_ld_32x32b_x32(acc + (uint32_t)c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const int col = c0 + 8 * ch;
            unsigned char* dst = tile1 + (col / a.D) * panel_bytes + swizzled_offset((uint32_t)r, (uint32_t)((col % a.D) * 2), span);
            *reinterpret_cast<uint4*>(dst) =
                make_uint4(pack_bf16x2(__uint_as_float(v[8 * ch]), __uint_as_float(v[8 * ch + 1])),
                           pack_bf16x2(__uint_as_float(v[8 * ch + 2]), __uint_as_float(v[8 * ch + 3])),
                           pack_bf16x2(__uint_as_float(v[8 * ch + 4]), __uint_as_float(v[8 * ch + 5])),
                           pack_bf16x2(__uint_as_float(v[8 * ch + 6]), __uint_as_float(v[8 * ch + 7])));
          }
        }
        acc_release();
        fence_proxy_async_smem();
        epi_bar_sync256();
        tr.ev(9, use);
        if (elected) {
          for (int pn = 0; pn * a.D < ncols; ++pn)
            tma_store_3d(&map_dz, tile1 + pn * panel_bytes, 0, (int)row0, (c * 256) / a.D + pn);
          tma_store_commit();
        }
      }
    }
    // bias gradients: one atomic per (thread, column) for the whole CTA
    const int j = et >> 1;
    if (a.g_post2_b != nullptr && 2 * j < Q) {
      if (b2acc[0] != 0.f) atomicAdd(a.g_post2_b + 2 * j, b2acc[0]);
      if (b2acc[1] != 0.f) atomicAdd(a.g_post2_b + 2 * j + 1, b2acc[1]);
    }
    if (a.g_post1_b != nullptr && 2 * j < P) {
      if (b1acc[0] != 0.f) atomicAdd(a.g_post1_b + 2 * j, b1acc[0]);
      if (b1acc[1] != 0.f) atomicAdd(a.g_post1_b + 2 * j + 1, b1acc[1]);
    }
    if (a.g_skip_b0 != nullptr && 2 * j < S) {
      if (bsacc[0] != 0.f) atomicAdd(a.g_skip_b0 + 2 * j, bsacc[0]);
      if (bsacc[1] != 0.f) atomicAdd(a.g_skip_b0 + 2 * j + 1, bsacc[1]);
    }
    if (elected) tma_store_wait_all<0>();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---- host side --------------------------------------------------------------------------------------
bool umma_post_supported(const wn_model* m) {
  static const bool disabled = getenv("WN_DISABLE_UMMA") != nullptr || getenv("WN_PREFER_CHAIN") != nullptr;
  const wn_arch& a = m->a;
  return !disabled && a.n_skip % 64 == 0 && a.n_post % 64 == 0 && a.n_skip <= 256 && a.n_post <= 256 &&
         a.n_quant == 256 && (a.n_dil == 16 || a.n_dil == 32 || a.n_dil == 64);
}

static int map2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint32_t box_inner,
                 uint32_t box_outer) {
  const uint64_t dims[2] = {inner, outer};
  const uint64_t strides[1] = {inner * 2};
  const uint32_t box[2] = {box_inner, box_outer};
  return make_tensor_map_bf16(out, base, 2, dims, strides, box, 128);
}

int launch_prep_umma(wn_model* m, const float* d_params, unsigned char* ws, cudaStream_t st) {
  const WorkspaceLayout& wl = m->wl;
  const wn_arch& a = m->a;
  k_prep_umma_weights<<<std::max(1, m->sm_count), 256, 0, st>>>(
      d_params, m->d_layers, m->L, a.n_dil, a.n_skip, a.n_post, a.n_quant, m->off_post1, m->off_post2,
      reinterpret_cast<bf16*>(ws + wl.wsT), reinterpret_cast<bf16*>(ws + wl.wsCat),
      reinterpret_cast<bf16*>(ws + wl.w1T), reinterpret_cast<bf16*>(ws + wl.w2T));
  WN_LAUNCH_CHECK();
  return WN_OK;
}

int launch_post_fwd_umma(wn_model* m, const float* d_params, unsigned char* ws, const int32_t* d_wav,
                         const int32_t* d_ids, int T, double* d_stats, float* d_logits, cudaStream_t st) {
  const WorkspaceLayout& wl = m->wl;
  const wn_arch& a = m->a;
  const int64_t rows = (int64_t)m->n_slots * T;
  const uint64_t LD = (uint64_t)m->L * a.n_dil, S = a.n_skip, P = a.n_post, Q = a.n_quant;
  CUtensorMap mz, mwsT, mw1T, mw2T, mh1, mh2, mdl;
  int rc;
  if ((rc = map2d(&mz, ws + wl.z, LD, (uint64_t)rows, UKB, UM))) return rc;
  static const bool single = getenv("WN_POST_SINGLE") != nullptr;  // A/B: one CTA per tile as in round 1
  const bool pair = !single && m->sm_count >= 2;
  // CTA pairs: every CTA loads half of the rows of a weight block
  const uint32_t bdiv = pair ? 2 : 1;
  if ((rc = map2d(&mwsT, ws + wl.wsT, LD, S, UKB, (uint32_t)S / bdiv))) return rc;
  if ((rc = map2d(&mw1T, ws + wl.w1T, S, P, UKB, (uint32_t)P / bdiv))) return rc;
  if ((rc = map2d(&mw2T, ws + wl.w2T, P, Q, UKB, (uint32_t)Q / bdiv))) return rc;
  if ((rc = map2d(&mh1, ws + wl.h1, S, (uint64_t)rows, UKB, UM))) return rc;
  if ((rc = map2d(&mh2, ws + wl.h2, P, (uint64_t)rows, UKB, UM))) return rc;
  if ((rc = map2d(&mdl, ws + wl.dlogits, Q, (uint64_t)rows, UKB, UM))) return rc;
  PostUmmaArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.trace = g_trace_layer == -2 ? g_trace_buf : nullptr;
  pa.params = d_params;
  pa.skip_bias = reinterpret_cast<const float*>(ws + wl.skip_bias);
  pa.off_post1_b = m->off_post1_b;
  pa.off_post2_b = m->off_post2_b;
  pa.wav = d_wav;
  pa.ids = d_ids;
  pa.stats = d_stats;
  pa.logits_out = d_logits;
  pa.T = T; pa.S = a.n_skip; pa.P = a.n_post; pa.Q = a.n_quant; pa.LD = (int)LD; pa.use_bias = a.use_bias;
  pa.rows = rows;
  pa.hm1 = reinterpret_cast<uint32_t*>(ws + wl.hm1);
  pa.hm2 = reinterpret_cast<uint32_t*>(ws + wl.hm2);
  if (pair) {
    // CTA pairs: every CTA loads half of the rows of a weight block
    const size_t smem = (size_t)4 * (UA_BYTES + UB_BYTES / 2) + UH_BYTES + 4096 + 1024;
    WN_CUDA_CHECK(cudaFuncSetAttribute(k_post_fwd_umma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope ps(PROF_POST_FWD, st);
    const int n_pairs = (int)((rows + 2 * UM - 1) / (2 * UM));
    int cap = m->sm_count;  // (test knob WN_PERSIST_GRID: many pairs per cluster)
    if (const char* e = getenv("WN_PERSIST_GRID")) if (atoi(e) > 0) cap = std::max(2, std::min(cap, atoi(e)));
    const int clusters = std::max(1, std::min(n_pairs, cap / 2));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(UPOST_P_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    WN_CUDA_CHECK(cudaLaunchKernelEx(&cfg, k_post_fwd_umma<true>, mz, mwsT, mw1T, mw2T, mh1, mh2, mdl, pa));
    WN_LAUNCH_CHECK();
    return WN_OK;
  }
  const size_t smem = (size_t)USTAGES * (UA_BYTES + UB_BYTES) + UH_BYTES + 4096 + 1024;
  WN_CUDA_CHECK(cudaFuncSetAttribute(k_post_fwd_umma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps(PROF_POST_FWD, st);
  const int n_tiles = (int)((rows + UM - 1) / UM);
  int grid = std::max(1, std::min(n_tiles, m->sm_count));
  if (const char* e = getenv("WN_PERSIST_GRID")) if (atoi(e) > 0) grid = std::max(1, std::min(grid, atoi(e)));
  k_post_fwd_umma<false><<<grid, UPOST_P_THREADS, smem, st>>>(mz, mwsT, mw1T, mw2T, mh1, mh2, mdl, pa);
  WN_LAUNCH_CHECK();
  return WN_OK;
}

}  // namespace wn

namespace wn {

// (one thread per (layer, column): as a loop over the layers in one thread this was L dependent global-memory round
// trips -- ~22 us for 30 layers on the step's critical path)
__global__ void k_bcast_skip_bias_umma(float* grads, const LayerDesc* layers, int L, int S) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x, l = 1 + blockIdx.y;
  if (s >= S || l >= L) return;
  grads[layers[l].skip_b + s] = grads[layers[0].skip_b + s];
}

int launch_post_bwd_umma(wn_model* m, unsigned char* ws, int T, float* d_grads, cudaStream_t st) {
  const WorkspaceLayout& wl = m->wl;
  const wn_arch& a = m->a;
  const int64_t rows = (int64_t)m->n_slots * T;
  const uint64_t LD = (uint64_t)m->L * a.n_dil, S = a.n_skip, P = a.n_post, Q = a.n_quant;
  const bf16* wbf = reinterpret_cast<const bf16*>(ws + wl.wbf);
  CUtensorMap mdl, mw2, mw1, mcat, mdp1, mdsk, mdz;
  int rc;
  if ((rc = map2d(&mdl, ws + wl.dlogits, Q, (uint64_t)rows, UKB, UM))) return rc;
  if ((rc = map2d(&mw2, wbf + m->off_post2, Q, P, UKB, (uint32_t)P))) return rc;   // POST2 [P][Q]: N = P, K = Q
  if ((rc = map2d(&mw1, wbf + m->off_post1, P, S, UKB, (uint32_t)S))) return rc;   // POST1 [S][P]: N = S, K = P
  if ((rc = map2d(&mcat, ws + wl.wsCat, S, LD, UKB, 256))) return rc;              // wsCat [LD][S]: N = LD, K = S
  if ((rc = map2d(&mdp1, ws + wl.dp1, P, (uint64_t)rows, UKB, UM))) return rc;
  if ((rc = map2d(&mdsk, ws + wl.dskip, S, (uint64_t)rows, UKB, UM))) return rc;
  {  // dz planes [L][B*T][D]: box = one layer's [128][D] panel
    const uint64_t D = a.n_dil;
    const uint64_t dims[3] = {D, (uint64_t)rows, (uint64_t)m->L};
    const uint64_t strides[2] = {D * 2, (uint64_t)rows * D * 2};
    const uint32_t box[3] = {(uint32_t)D, UM, 1};
    if ((rc = make_tensor_map_bf16(&mdz, ws + wl.dz, 3, dims, strides, box, (int)D * 2))) return rc;
  }
  PostBwdUmmaArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.trace = g_trace_layer == -2 ? g_trace_buf : nullptr;
  pa.h1 = reinterpret_cast<const bf16*>(ws + wl.h1);
  pa.h2 = reinterpret_cast<const bf16*>(ws + wl.h2);
  pa.hm1 = reinterpret_cast<const uint32_t*>(ws + wl.hm1);
  pa.hm2 = reinterpret_cast<const uint32_t*>(ws + wl.hm2);
  if (a.use_bias) {
    pa.g_post2_b = d_grads + m->off_post2_b;
    pa.g_post1_b = d_grads + m->off_post1_b;
    pa.g_skip_b0 = d_grads + m->layers[0].skip_b;
  }
  pa.S = a.n_skip; pa.P = a.n_post; pa.Q = a.n_quant; pa.LD = (int)LD; pa.D = a.n_dil; pa.rows = rows;
  const size_t smem = (size_t)2 * UH_BYTES + (size_t)UBSTAGES * UB_BYTES + 1024;
  WN_CUDA_CHECK(cudaFuncSetAttribute(k_post_bwd_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  {
    ProfScope ps(PROF_POST_BWD, st);
    const int n_tiles = (int)((rows + UM - 1) / UM);
    int grid = std::max(1, std::min(n_tiles, m->sm_count));
    if (const char* e = getenv("WN_PERSIST_GRID")) if (atoi(e) > 0) grid = std::max(1, std::min(grid, atoi(e)));
    k_post_bwd_umma<<<grid, UPOST_P_THREADS, smem, st>>>(mdl, mw2, mw1, mcat, mdp1, mdsk, mdz, pa);
    WN_LAUNCH_CHECK();
  }
  if (a.use_bias && m->L > 1) {
    k_bcast_skip_bias_umma<<<dim3((a.n_skip + 127) / 128, m->L - 1), 128, 0, st>>>(d_grads, m->d_layers, m->L, a.n_skip);
    WN_LAUNCH_CHECK();
  }
  return WN_OK;
}

}  // namespace wn

// =====================================================================================================
// k_gemm_umma: the post-net as a chain of plain tcgen05 GEMMs, for the shapes the fused kernels above do not hold in
// shared memory / tensor memory at once (n_skip or n_post = 512: the reference's own par/arch*.json all use 512).
//     out[rows x N] = epilogue( A[rows x K] . B[N x K]^T ),  N processed in chunks of <= 256 columns
// Persistent, one CTA per SM; work items = (128-row tile, N chunk), chunks innermost so the A tile's re-read hits L2.
// warp 0 TMA producer (3-stage ring of A [128 x 64] + B [<=256 x 64] K blocks), warp 1 MMA issuer, 8 epilogue warps
// (thread <-> (row, column half)); two 256-column TMEM accumulators alternate between items, so item j+1's
// contraction overlaps item j's epilogue.  Epilogues (mode):
//   0  bf16(relu(acc + bias))                      -> staging tile -> TMA store        (h1, h2: tmodel.py:187-215)
//   1  bf16(acc * (H > 0))                         -> staging tile -> TMA store        (dp1, dskip)
//   2  bf16(acc) as per-layer [128 x D] panels     -> TMA stores into the dz planes
//   3  logits = acc + bias; masked softmax cross entropy, statistics, dlogits          (tmodel.py:228-249)
//   7  bf16(acc)                                   -> staging tile -> TMA store        (local-conditioning chain)
// =====================================================================================================
namespace wn {

constexpr int GEMM_STAGES = 3;

struct GemmUmmaArgs {
  int64_t rows;
  int K, N, mode, D;
  const float* bias;   // modes 0, 3: [N] fp32 or nullptr
  const bf16* H;       // mode 1: [rows][N] relu outputs (the mask)
  const int32_t* wav;  // mode 3
  const int32_t* ids;
  double* stats;
  float* logits_out;
  int T;
  // ---- per-layer use (modes 4, 5): tiles are (slot, 128 timesteps); A and the output are 3-D maps [cols][rows][slot]
  int tps;            // tiles per slot (0: flat rows, 2-D maps)
  int a_col0;         // first A column (the z stash holds every layer side by side)
  int a_k_split;      // K elements that come from A rows t0 + ...; the rest from rows t0 + a_row_off2 (second conv tap)
  int a_row_off2;
  int out_col0, out_row_off;
  const float* bias2; // mode 4: GATE bias (bias = SIGNAL bias); both [D]
  const bf16* X;      // mode 5: the layer input x[t] for the residual add: X[(slot * x_slot_rows + x_row_off + t) * N + c]
  int x_slot_rows, x_row_off;
  // modes 5 (3-D), 6: column sums of the OUTPUT tile, accumulated over the CTA's tiles (bias gradients): columns
  // [0, colsum_split) -> colsum_out, the rest -> colsum_out2
  float* colsum_out;
  float* colsum_out2;
  int colsum_split;
  // b_resident: the whole B operand (N <= 256, K * N * 2 bytes) is loaded into shared memory once per CTA and only A
  // goes through the ring (a_stages deep).  Re-streaming B from L2 for every 128-row tile is what bounds the ring
  // version when B is larger than the A tile (wide layers: 128 KB of conv weights against a 64 KB activation tile).
  int b_resident, a_stages;
  int l2_prefetch;  // producer prefetches the next tile's A blocks into L2 (3-D maps only)
  // a_planes: A is a stack of K / 64 planes [plane][rows][64] (the per-layer dcond planes): K block kb of a row tile is
  // the 3-D box (0, row0, kb) of map_a
  int a_planes;
};
constexpr int GEMM_MAX_STAGES = 8;

__global__ void __launch_bounds__(UPOST_P_THREADS, 1)
k_gemm_umma(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
            const __grid_constant__ CUtensorMap map_out, GemmUmmaArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nst = a.b_resident ? a.a_stages : GEMM_STAGES;
  const int b_blk = a.b_resident ? min(256, a.N) * 128 : UB_BYTES;  // bytes of one K block of B
  unsigned char* stage_a = smem;                                   // nst x 16 KB
  unsigned char* stage_b = smem + nst * UA_BYTES;                  // ring: GEMM_STAGES x 32 KB; resident: K / 64 blocks
  unsigned char* otile = stage_b + (a.b_resident ? (a.K / UKB) * b_blk : GEMM_STAGES * UB_BYTES);  // staging tile
  __shared__ __align__(8) uint64_t full_bar[GEMM_MAX_STAGES], empty_bar[GEMM_MAX_STAGES], acc_full[2], acc_empty[2], b_full;
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[512];
  __shared__ float x_mx[2][128], x_sum[2][128], x_vl[2][128];
  __shared__ int x_arg[2][128];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tiles = a.tps > 0 ? a.tps * (int)(a.rows / a.T) : (int)((a.rows + UM - 1) / UM);
  const int n_my = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int nchunk = (a.N + 255) / 256, nkb = a.K / UKB;
  const int box_n = min(256, a.N);  // rows of B one TMA box brings (a partial last chunk is zero filled)
  pdl_launch_dependents();  // the next launch of the chain may set up as soon as this CTA releases its resources
  if (tid == 0) {
    for (int i = 0; i < GEMM_MAX_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 256);
    }
    mbar_init(&b_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_out);
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  pdl_wait();  // everything below reads or writes global memory the previous kernels of the stream own
  if (a.mode == 4 || a.mode == 6) {  // SIGNAL_BIAS | 0.5 * GATE_BIAS (the GATE filter copy is pre-scaled by 0.5: sigmoid via tanh)
    const int Dh = a.N / 2;
    for (int i = tid; i < 512; i += UPOST_P_THREADS)
      bias_s[i] = i < Dh ? (a.bias != nullptr ? a.bias[i] : 0.f)
                : i < 2 * Dh ? (a.bias2 != nullptr ? 0.5f * a.bias2[i - Dh] : 0.f) : 0.f;
  } else {
    for (int i = tid; i < 512; i += UPOST_P_THREADS) bias_s[i] = (a.bias != nullptr && i < a.N) ? a.bias[i] : 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      if (a.b_resident && n_my > 0) {  // the weights do not depend on the previous kernel's output: no need to order
        mbar_expect_tx(&b_full, (uint32_t)(nkb * b_blk));
        for (int kb = 0; kb < nkb; ++kb) tma_load_2d(stage_b + kb * b_blk, &map_b, &b_full, kb * UKB, 0);
      }
      int st = 0;
      uint32_t ph = 1;
      // L2 prefetch of the next tile's A blocks (per-layer launches): with B resident only a_stages x 16 KB of A are in
      // flight per SM, too little to cover DRAM latency; the ring then reads at L2 latency
      auto prefetch_tile = [&](int i_tile) {
        const int tile2 = (int)blockIdx.x + i_tile * (int)gridDim.x;
        const int sb2 = tile2 / a.tps, t02 = (tile2 % a.tps) * UM;
        for (int kb = 0; kb < nkb; ++kb) {
          int k0 = kb * UKB, roff = 0;
          if (a.a_k_split > 0 && k0 >= a.a_k_split) { k0 -= a.a_k_split; roff = a.a_row_off2; }
          tma_prefetch_l2_3d(&map_a, a.a_col0 + k0, t02 + roff, sb2);
        }
      };
      const bool pf = a.l2_prefetch && a.tps > 0;
      if (pf && n_my > 1) prefetch_tile(1);
      for (int i = 0; i < n_my; ++i) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int row0 = tile * UM;
        const int sb = a.tps > 0 ? tile / a.tps : 0, t0 = a.tps > 0 ? (tile % a.tps) * UM : 0;
        if (pf && i >= 1 && i + 1 < n_my) prefetch_tile(i + 1);
        for (int c = 0; c < nchunk; ++c)
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&empty_bar[st], ph);
            mbar_expect_tx(&full_bar[st], (uint32_t)(UA_BYTES + (a.b_resident ? 0 : box_n * 128)));
            if (a.tps > 0) {
              int k0 = kb * UKB, roff = 0;
              if (a.a_k_split > 0 && k0 >= a.a_k_split) { k0 -= a.a_k_split; roff = a.a_row_off2; }
              tma_load_3d(stage_a + st * UA_BYTES, &map_a, &full_bar[st], a.a_col0 + k0, t0 + roff, sb);
            } else if (a.a_planes) {
              tma_load_3d(stage_a + st * UA_BYTES, &map_a, &full_bar[st], 0, row0, kb);
            } else {
              tma_load_2d(stage_a + st * UA_BYTES, &map_a, &full_bar[st], kb * UKB, row0);
            }
            if (!a.b_resident) tma_load_2d(stage_b + st * UB_BYTES, &map_b, &full_bar[st], kb * UKB, c * 256);
            if (++st == nst) { st = 0; ph ^= 1u; }
          }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int item = 0, st = 0;
      uint32_t ph = 0;
      if (a.b_resident && n_my > 0) mbar_wait(&b_full, 0);
      for (int i = 0; i < n_my; ++i)
        for (int c = 0; c < nchunk; ++c, ++item) {
          const int buf = item & 1;
          const int w = min(256, a.N - c * 256);
          const uint32_t idesc = make_idesc_bf16(UM, (w + 15) & ~15);
          mbar_wait(&acc_empty[buf], ((uint32_t)(item >> 1) & 1u) ^ 1u);
          tc_fence_after_sync();
          const uint32_t acc = tmem_base + (uint32_t)buf * 256;
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&full_bar[st], ph);
            tc_fence_after_sync();
            const uint32_t sa = smem_u32(stage_a + st * UA_BYTES),
                           sb = smem_u32(a.b_resident ? stage_b + kb * b_blk : stage_b + st * UB_BYTES);
#pragma unroll
            for (int k = 0; k < UKB / 16; ++k)
              mma_bf16_ss(acc, make_kmajor_desc(sa, 128, k * 32), make_kmajor_desc(sb, 128, k * 32), idesc, (kb | k) != 0);
            mma_commit(&empty_bar[st]);
            if (++st == nst) { st = 0; ph ^= 1u; }
          }
          mma_commit(&acc_full[buf]);
        }
    }
  } else {
    const int e = warp - 2, q4 = warp & 3, half = e >> 2;
    const int r = q4 * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const bool elected = (warp == 2 && lane == 0);
    uint32_t v[32];
    uint32_t pk[16];
    float acc_x = 0.f, acc_n = 0.f, acc_d = 0.f;
    float cs_acc[2] = {0.f, 0.f};
    const int et = e * 32 + lane;  // 0..255
    // Modes 5 and 6 add a per-row global operand (x[t] / the dz plane) in the epilogue.  Loading it after the accumulator
    // is ready exposes one DRAM latency per 32-column step (ncu: the 128 x 128 residual GEMM took longer than the
    // 128 x 256 x 256 conv GEMM).  When this thread's slice of the row is <= 64 columns it is prefetched into xq one
    // tile ahead: the loads of tile i + 1 are issued right after tile i's TMEM reads and fly during its store phase.
    const int x_span = a.mode == 5 ? a.N / 2 : a.mode == 6 ? a.N / 4 : 0;  // columns of the row this thread consumes
    const bool x_pre = a.X != nullptr && nchunk == 1 && x_span > 0 && x_span <= 64;
    uint32_t xq[32];
    auto load_xq = [&](int i_tile) {
      const int tile2 = (int)blockIdx.x + i_tile * (int)gridDim.x;
      const int sb2 = a.tps > 0 ? tile2 / a.tps : 0, t02 = a.tps > 0 ? (tile2 % a.tps) * UM : 0;
      const bf16* src;
      bool ok;
      if (a.mode == 5) {
        ok = a.tps > 0 ? t02 + r < a.T : (int64_t)tile2 * UM + r < a.rows;
        src = (a.tps > 0 ? a.X + ((size_t)sb2 * a.x_slot_rows + a.x_row_off + t02 + r) * a.N
                         : a.X + ((size_t)tile2 * UM + r) * a.N) + half * x_span;
      } else {
        ok = t02 + r < a.T;
        src = a.X + ((size_t)sb2 * a.T + t02 + r) * (a.N / 2) + half * x_span;
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (ok && u * 32 < x_span) {
          const uint4* x4 = reinterpret_cast<const uint4*>(src + u * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 t4 = __ldg(x4 + q);
            xq[16 * u + 4 * q] = t4.x; xq[16 * u + 4 * q + 1] = t4.y; xq[16 * u + 4 * q + 2] = t4.z; xq[16 * u + 4 * q + 3] = t4.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) xq[16 * u + j] = 0u;
        }
      }
    };
    if (x_pre && n_my > 0) load_xq(0);
    int item = 0;
    for (int i = 0; i < n_my; ++i) {
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const int64_t row0 = (int64_t)tile * UM;
      const int64_t row = row0 + r;
      const bool in_range = row < a.rows;
      const int sb = a.tps > 0 ? tile / a.tps : 0, t0 = a.tps > 0 ? (tile % a.tps) * UM : 0;
      for (int c = 0; c < nchunk; ++c, ++item) {
        const int buf = item & 1;
        const int n0 = c * 256, w = min(256, a.N - n0);
        mbar_wait(&acc_full[buf], (uint32_t)(item >> 1) & 1u);
        tc_fence_after_sync();
        const uint32_t acc = tmem_base + (uint32_t)buf * 256 + lane_sel;
        if (a.mode != 3) {
          if (elected) tma_store_wait_read<0>();  // the previous item's stores have finished reading the staging tile
          epi_bar_sync256();
        }
        if (a.mode == 0 || a.mode == 1 || a.mode == 7) {
          const int cb = half * (w / 2);
          for (int c0 = cb; c0 < cb + w / 2; c0 += 32) {
            tmem_ld_32x32b_x32(acc + (uint32_t)c0, v);
            tmem_ld_wait();
            if (a.mode == 0) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                pk[j] = pack_bf16x2(fmaxf(__uint_as_float(v[2 * j]) + bias_s[n0 + c0 + 2 * j], 0.f),
                                    fmaxf(__uint_as_float(v[2 * j + 1]) + bias_s[n0 + c0 + 2 * j + 1], 0.f));
            } else if (a.mode == 7) {
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
            } else if (in_range) {
              mask_pack32(v, a.H + (size_t)row * a.N + n0 + c0, pk);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = 0u;
            }
            htile_store32(otile, r, c0, pk);
          }
          tc_fence_before_sync();
          mbar_arrive(&acc_empty[buf]);
          fence_proxy_async_smem();
          epi_bar_sync256();
          if (elected) {
            for (int kb = 0; kb < w / UKB; ++kb) tma_store_2d(&map_out, otile + kb * UA_BYTES, n0 + kb * UKB, (int)row0);
            tma_store_commit();
          }
        } else if (a.mode == 4) {
          // gate: columns [h*D, h*D + D/2) = SIGNAL channels [h*D/2, ...), [h*D + D/2, (h+1)*D) = the same GATE channels
          // (weight rows permuted by k_prep_wide_weights); z = tanh(s) * (0.5 tanh(g/2) + 0.5)   (tmodel.py:167)
          const int Dn = a.N / 2, Dh = Dn / 2;  // z channels, channels per column half
          uint32_t g[32];
          for (int c0 = 0; c0 < Dh; c0 += 32) {
            tmem_ld_32x32b_x32(acc + (uint32_t)(half * Dn + c0), v);
            tmem_ld_32x32b_x32(acc + (uint32_t)(half * Dn + Dh + c0), g);
            tmem_ld_wait();
            const int ch = half * Dh + c0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float z0 = tanh_fast(__uint_as_float(v[2 * j]) + bias_s[ch + 2 * j]) *
                               fmaf(0.5f, tanh_fast(__uint_as_float(g[2 * j]) + bias_s[Dn + ch + 2 * j]), 0.5f);
              const float z1 = tanh_fast(__uint_as_float(v[2 * j + 1]) + bias_s[ch + 2 * j + 1]) *
                               fmaf(0.5f, tanh_fast(__uint_as_float(g[2 * j + 1]) + bias_s[Dn + ch + 2 * j + 1]), 0.5f);
              pk[j] = pack_bf16x2(z0, z1);
            }
            htile_store32(otile, r, ch, pk);
          }
          tc_fence_before_sync();
          mbar_arrive(&acc_empty[buf]);
          fence_proxy_async_smem();
          epi_bar_sync256();
          if (elected) {
            for (int kb = 0; kb < Dn / UKB; ++kb)
              tma_store_3d(&map_out, otile + kb * UA_BYTES, a.out_col0 + kb * UKB, a.out_row_off + t0, sb);
            tma_store_commit();
          }
        } else if (a.mode == 5) {
          // residual: x' = bf16(x[t] + z . RESIDUAL + bias) written into the next layer's input  (tmodel.py:325)
          const bool row_ok = a.X != nullptr && (a.tps > 0 ? t0 + r < a.T : in_range);
          const bf16* xrow = a.tps > 0 ? a.X + ((size_t)sb * a.x_slot_rows + a.x_row_off + t0 + r) * a.N
                                       : a.X + (size_t)row * a.N;
          const int cb = half * (w / 2);
          for (int c0 = cb; c0 < cb + w / 2; c0 += 32) {
            tmem_ld_32x32b_x32(acc + (uint32_t)c0, v);
            tmem_ld_wait();
            uint32_t xw[16];
            if (x_pre) {
              const bool second = c0 != cb;
#pragma unroll
              for (int j = 0; j < 16; ++j) xw[j] = second ? xq[16 + j] : xq[j];
            } else if (row_ok) {
              const uint4* x4 = reinterpret_cast<const uint4*>(xrow + c0);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 t4 = __ldg(x4 + q);
                xw[4 * q] = t4.x; xw[4 * q + 1] = t4.y; xw[4 * q + 2] = t4.z; xw[4 * q + 3] = t4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) xw[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j)
              pk[j] = pack_bf16x2(__uint_as_float(v[2 * j]) + __uint_as_float(xw[j] << 16) + bias_s[c0 + 2 * j],
                                  __uint_as_float(v[2 * j + 1]) + __uint_as_float(xw[j] & 0xffff0000u) + bias_s[c0 + 2 * j + 1]);
            htile_store32(otile, r, c0, pk);
          }
          if (x_pre && i + 1 < n_my) load_xq(i + 1);
          tc_fence_before_sync();
          mbar_arrive(&acc_empty[buf]);
          fence_proxy_async_smem();
          epi_bar_sync256();
          if (elected) {
            for (int kb = 0; kb < w / UKB; ++kb) {
              if (a.tps > 0) tma_store_3d(&map_out, otile + kb * UA_BYTES, a.out_col0 + kb * UKB, a.out_row_off + t0, sb);
              else tma_store_2d(&map_out, otile + kb * UA_BYTES, a.out_col0 + kb * UKB, (int)row0);
            }
            tma_store_commit();
          }
          if (a.colsum_out != nullptr && a.tps > 0) tile_colsum_acc(otile, w, et, min(UM, a.T - t0), cs_acc);
        } else if (a.mode == 6) {
          // gate backward: recompute th, sg from the pre-activations; dz (skip + residual part) from the dz plane;
          // dv_s = dz sg (1 - th^2), dv_g = dz th sg (1 - sg) -> dv [.. x 2D] (SIGNAL | GATE)
          const int Dn = a.N / 2, Dh = Dn / 2;
          const bool row_ok = t0 + r < a.T;
          const bf16* dzrow = a.X + ((size_t)sb * a.T + t0 + r) * Dn;
          uint32_t g[32], pg[16];
          for (int c0 = 0; c0 < Dh; c0 += 32) {
            tmem_ld_32x32b_x32(acc + (uint32_t)(half * Dn + c0), v);
            tmem_ld_32x32b_x32(acc + (uint32_t)(half * Dn + Dh + c0), g);
            const int ch = half * Dh + c0;
            uint32_t dzw[16];
            if (x_pre) {
              const bool second = c0 != 0;
#pragma unroll
              for (int j = 0; j < 16; ++j) dzw[j] = second ? xq[16 + j] : xq[j];
            } else if (row_ok) {
              const uint4* d4 = reinterpret_cast<const uint4*>(dzrow + ch);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 t4 = __ldg(d4 + q);
                dzw[4 * q] = t4.x; dzw[4 * q + 1] = t4.y; dzw[4 * q + 2] = t4.z; dzw[4 * q + 3] = t4.w;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) dzw[j] = 0u;
            }
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float ds[2], dg[2];
#pragma unroll
              for (int e2 = 0; e2 < 2; ++e2) {
                const float th = tanh_fast(__uint_as_float(v[2 * j + e2]) + bias_s[ch + 2 * j + e2]);
                const float u = tanh_fast(__uint_as_float(g[2 * j + e2]) + bias_s[Dn + ch + 2 * j + e2]);
                const float sg = fmaf(0.5f, u, 0.5f);
                const float dz = e2 == 0 ? __uint_as_float(dzw[j] << 16) : __uint_as_float(dzw[j] & 0xffff0000u);
                ds[e2] = (dz * sg) * fmaf(-th, th, 1.f);
                dg[e2] = (dz * th) * (sg * (1.f - sg));
              }
              pk[j] = pack_bf16x2(ds[0], ds[1]);
              pg[j] = pack_bf16x2(dg[0], dg[1]);
            }
            htile_store32(otile, r, ch, pk);
            htile_store32(otile, r, Dn + ch, pg);
          }
          if (x_pre && i + 1 < n_my) load_xq(i + 1);
          tc_fence_before_sync();
          mbar_arrive(&acc_empty[buf]);
          fence_proxy_async_smem();
          epi_bar_sync256();
          if (elected) {
            for (int kb = 0; kb < a.N / UKB; ++kb)
              tma_store_3d(&map_out, otile + kb * UA_BYTES, kb * UKB, t0, sb);
            tma_store_commit();
          }
          if (a.colsum_out != nullptr) tile_colsum_acc(otile, a.N, et, min(UM, a.T - t0), cs_acc);
        } else if (a.mode == 2) {
          // panels of PW = min(D, 64) columns (one swizzle span per row); a plane wider than 64 columns is two panels
          const int PW = min(a.D, 64), span = 2 * PW, panel_bytes = UM * span;
          const int cb = half * 128, ce = min(w, cb + 128);
          for (int c0 = cb; c0 < ce; c0 += 32) {
            tmem_ld_32x32b_x32(acc + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              const int col = c0 + 8 * ch;
              unsigned char* dst = otile + (col / PW) * panel_bytes + swizzled_offset((uint32_t)r, (uint32_t)((col % PW) * 2), span);
              *reinterpret_cast<uint4*>(dst) =
                  make_uint4(pack_bf16x2(__uint_as_float(v[8 * ch]), __uint_as_float(v[8 * ch + 1])),
                             pack_bf16x2(__uint_as_float(v[8 * ch + 2]), __uint_as_float(v[8 * ch + 3])),
                             pack_bf16x2(__uint_as_float(v[8 * ch + 4]), __uint_as_float(v[8 * ch + 5])),
                             pack_bf16x2(__uint_as_float(v[8 * ch + 6]), __uint_as_float(v[8 * ch + 7])));
            }
          }
          tc_fence_before_sync();
          mbar_arrive(&acc_empty[buf]);
          fence_proxy_async_smem();
          epi_bar_sync256();
          if (elected) {
            for (int pn = 0; pn * PW < w; ++pn) {
              const int col = n0 + pn * PW;
              tma_store_3d(&map_out, otile + pn * panel_bytes, col % a.D, (int)row0, col / a.D);
            }
            tma_store_commit();
          }
        } else {
          // ---- mode 3: logits, masked softmax cross entropy, dlogits (N == 256, one chunk per tile) ----
          const int Q = a.N;
          const int b = in_range ? (int)(row / a.T) : 0, t = in_range ? (int)(row % a.T) : 0;
          const bool has_next = in_range && (t + 1 < a.T);
          const bool valid = has_next && (a.ids[(size_t)b * a.T + t + 1] != 0);  // tmodel.py:232
          const int label_raw = has_next ? a.wav[(size_t)b * a.T + t + 1] : 0;    // tmodel.py:230
          const bool label_ok = label_raw >= 0 && label_raw < Q;                  // out of range: all-zero one-hot row
          const int label = label_ok ? label_raw : -1;
          const int cb = half * (Q / 2);
          float mx = -INFINITY, sum = 0.f, vl = 0.f;
          int arg = 0;
          for (int c0 = cb; c0 < cb + Q / 2; c0 += 32) {
            tmem_ld_32x32b_x32(acc + (uint32_t)c0, v);
            tmem_ld_wait();
            float cm = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float x = __uint_as_float(v[j]) + bias_s[c0 + j];
              v[j] = __float_as_uint(x);
              if (x > mx && x > cm) arg = c0 + j;
              cm = fmaxf(cm, x);
              if (c0 + j == label) vl = x;
              if (a.logits_out != nullptr && in_range) a.logits_out[(size_t)row * Q + c0 + j] = x;
            }
            const float nm = fmaxf(mx, cm);
            float cs = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) cs += __expf(__uint_as_float(v[j]) - nm);
            sum = sum * __expf(mx - nm) + cs;
            mx = nm;
          }
          x_mx[half][r] = mx;
          x_sum[half][r] = sum;
          x_arg[half][r] = arg;
          x_vl[half][r] = vl;
          if (elected) tma_store_wait_read<0>();
          epi_bar_sync256();
          const float om = x_mx[half ^ 1][r], os = x_sum[half ^ 1][r];
          const float M = fmaxf(mx, om);
          const float tot = sum * __expf(mx - M) + os * __expf(om - M);
          const float inv = 1.f / tot;
          for (int c0 = cb; c0 < cb + Q / 2; c0 += 32) {
            tmem_ld_32x32b_x32(acc + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float x0 = __uint_as_float(v[2 * j]) + bias_s[c0 + 2 * j];
              const float x1 = __uint_as_float(v[2 * j + 1]) + bias_s[c0 + 2 * j + 1];
              const float g0 = __expf(x0 - M) * inv - ((c0 + 2 * j) == label ? 1.f : 0.f);
              const float g1 = __expf(x1 - M) * inv - ((c0 + 2 * j + 1) == label ? 1.f : 0.f);
              pk[j] = valid ? pack_bf16x2(g0, g1) : 0u;
            }
            htile_store32(otile, r, c0, pk);
          }
          tc_fence_before_sync();
          mbar_arrive(&acc_empty[buf]);
          fence_proxy_async_smem();
          epi_bar_sync256();
          if (elected) {
            for (int kb = 0; kb < Q / UKB; ++kb) tma_store_2d(&map_out, otile + kb * UA_BYTES, kb * UKB, (int)row0);
            tma_store_commit();
          }
          if (half == 0 && valid) {
            const int garg = (x_mx[1][r] > x_mx[0][r]) ? x_arg[1][r] : x_arg[0][r];
            const float gvl = label < Q / 2 ? x_vl[0][r] : x_vl[1][r];
            acc_x += label_ok ? __logf(tot) + M - gvl : 0.f;
            acc_n += 1.f;
            acc_d += fabsf((float)((label_ok ? label : 0) - garg));
          }
          epi_bar_sync256();  // the partials of this tile are consumed before the next tile overwrites them
        }
      }
    }
    if (a.colsum_out != nullptr && (a.mode == 6 || a.mode == 5) && 2 * (et >> 1) < a.N) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int col = 2 * (et >> 1) + q;
        float* dst = col < a.colsum_split ? a.colsum_out + col : a.colsum_out2 + (col - a.colsum_split);
        if (cs_acc[q] != 0.f) atomicAdd(dst, cs_acc[q]);
      }
    }
    if (a.mode == 3) {
      acc_x = warp_sum(acc_x);
      acc_n = warp_sum(acc_n);
      acc_d = warp_sum(acc_d);
      if (lane == 0 && acc_n != 0.f) {
        atomicAdd(a.stats + WN_STAT_XENT_SUM, (double)acc_x);
        atomicAdd(a.stats + WN_STAT_N_VALID, (double)acc_n);
        atomicAdd(a.stats + WN_STAT_DIFF_SUM, (double)acc_d);
      }
    }
    if (elected) tma_store_wait_all<0>();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// column sums of a bf16 [rows x N] matrix, added to out[N] (bias gradients on the GEMM-chain path)
__global__ void k_colsum_bf16(const bf16* __restrict__ x, int64_t rows, int ld, int N, float* __restrict__ out) {
  const int c2 = blockIdx.x * blockDim.x + threadIdx.x;  // column pair
  if (2 * c2 >= N) return;
  const int64_t r0 = (int64_t)blockIdx.y * ((rows + gridDim.y - 1) / gridDim.y);
  const int64_t r1 = min(rows, r0 + (rows + gridDim.y - 1) / gridDim.y);
  float s0 = 0.f, s1 = 0.f;
  for (int64_t r = r0; r < r1; ++r) {
    const uint32_t wv = *reinterpret_cast<const uint32_t*>(x + r * ld + 2 * c2);
    s0 += __uint_as_float(wv << 16);
    s1 += __uint_as_float(wv & 0xffff0000u);
  }
  if (s0 != 0.f) atomicAdd(out + 2 * c2, s0);
  if (s1 != 0.f) atomicAdd(out + 2 * c2 + 1, s1);
}

// shapes the GEMM chain covers when the fused kernels do not
bool umma_post_chain_supported(const wn_model* m) {
  static const bool disabled = getenv("WN_DISABLE_UMMA") != nullptr || getenv("WN_DISABLE_UMMA_CHAIN") != nullptr;
  const wn_arch& a = m->a;
  auto ok = [](int n) { return n % 64 == 0 && (n <= 256 || n % 256 == 0) && n <= 512; };
  return !disabled && ok(a.n_skip) && ok(a.n_post) && a.n_quant == 256 && ((int64_t)m->L * a.n_dil) % 64 == 0 &&
         (a.n_dil == 16 || a.n_dil == 32 || a.n_dil == 64 || a.n_dil == 128);
}

// shared-memory plan of k_gemm_umma: resident B when it, the staging tile and >= 3 A stages fit (WN_GEMM_NO_RESIDENT=1:
// always the ring)
static size_t gemm_smem_plan(GemmUmmaArgs& ga) {
  static const bool off = getenv("WN_GEMM_NO_RESIDENT") != nullptr;
  ga.b_resident = 0;
  ga.a_stages = GEMM_STAGES;
  const size_t ring = (size_t)GEMM_STAGES * (UA_BYTES + UB_BYTES) + UH_BYTES + 1024;
  if (off || ga.N > 256 || ga.K % UKB != 0) return ring;
  const size_t b_bytes = (size_t)(ga.K / UKB) * std::min(256, ga.N) * 128;
  const size_t o_bytes = (size_t)UM * (ga.mode == 4 ? ga.N / 2 : ga.N) * 2;
  const size_t budget = 220 * 1024;
  if (1024 + b_bytes + o_bytes + 3 * (size_t)UA_BYTES > budget) return ring;
  const int stages = (int)std::min<size_t>(GEMM_MAX_STAGES, (budget - 1024 - b_bytes - o_bytes) / UA_BYTES);
  static const bool pf = getenv("WN_GEMM_PREFETCH") != nullptr;  // measured: no gain (7.58 vs 7.53 ms wide step), off by default
  ga.b_resident = 1;
  ga.a_stages = stages;
  ga.l2_prefetch = pf && stages < 6;
  return 1024 + (size_t)stages * UA_BYTES + b_bytes + o_bytes;
}

static int launch_gemm_umma(wn_model* m, const void* A, int K, const void* B, int N, void* out, int64_t rows,
                            GemmUmmaArgs ga, cudaStream_t st) {
  CUtensorMap ma, mb, mo;
  int rc;
  if (ga.a_planes) {
    const uint64_t dims[3] = {UKB, (uint64_t)rows, (uint64_t)(K / UKB)};
    const uint64_t strides[2] = {UKB * 2, (uint64_t)rows * UKB * 2};
    const uint32_t box[3] = {UKB, UM, 1};
    if ((rc = make_tensor_map_bf16(&ma, A, 3, dims, strides, box, 128))) return rc;
  } else if ((rc = map2d(&ma, A, (uint64_t)K, (uint64_t)rows, UKB, UM))) {
    return rc;
  }
  if ((rc = map2d(&mb, B, (uint64_t)K, (uint64_t)N, UKB, (uint32_t)std::min(256, N)))) return rc;
  if (ga.mode == 2) {
    const uint64_t D = ga.D, PW = std::min<uint64_t>(D, 64);
    const uint64_t dims[3] = {D, (uint64_t)rows, (uint64_t)(N / ga.D)};
    const uint64_t strides[2] = {D * 2, (uint64_t)rows * D * 2};
    const uint32_t box[3] = {(uint32_t)PW, UM, 1};
    if ((rc = make_tensor_map_bf16(&mo, out, 3, dims, strides, box, (int)PW * 2))) return rc;
  } else {
    if ((rc = map2d(&mo, out, (uint64_t)N, (uint64_t)rows, UKB, UM))) return rc;
  }
  ga.rows = rows; ga.K = K; ga.N = N;
  const size_t smem = gemm_smem_plan(ga);
  WN_CUDA_CHECK(cudaFuncSetAttribute(k_gemm_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  const int n_tiles = (int)((rows + UM - 1) / UM);
  int grid = std::max(1, std::min(n_tiles, m->sm_count));
  if (const char* e = getenv("WN_PERSIST_GRID")) if (atoi(e) > 0) grid = std::max(1, std::min(grid, atoi(e)));
  WN_CUDA_CHECK(launch_pdl(k_gemm_umma, grid, UPOST_P_THREADS, smem, st, ma, mb, mo, ga));
  WN_LAUNCH_CHECK();
  return WN_OK;
}

// skip sum + post-net + loss as three GEMMs (h1, h2 are stashed for the backward exactly as the fused kernel does)
int launch_post_fwd_chain_umma(wn_model* m, const float* d_params, unsigned char* ws, const int32_t* d_wav,
                               const int32_t* d_ids, int T, double* d_stats, float* d_logits, cudaStream_t st) {
  const WorkspaceLayout& wl = m->wl;
  const wn_arch& a = m->a;
  const int64_t rows = (int64_t)m->n_slots * T;
  const int LD = m->L * a.n_dil, S = a.n_skip, P = a.n_post, Q = a.n_quant;
  ProfScope ps(PROF_POST_FWD, st);
  GemmUmmaArgs ga;
  int rc;
  memset(&ga, 0, sizeof(ga));
  ga.mode = 0;
  ga.bias = a.use_bias ? reinterpret_cast<const float*>(ws + wl.skip_bias) : nullptr;
  if ((rc = launch_gemm_umma(m, ws + wl.z, LD, ws + wl.wsT, S, ws + wl.h1, rows, ga, st))) return rc;
  ga.bias = a.use_bias ? d_params + m->off_post1_b : nullptr;
  if ((rc = launch_gemm_umma(m, ws + wl.h1, S, ws + wl.w1T, P, ws + wl.h2, rows, ga, st))) return rc;
  ga.mode = 3;
  ga.bias = a.use_bias ? d_params + m->off_post2_b : nullptr;
  ga.wav = d_wav; ga.ids = d_ids; ga.stats = d_stats; ga.logits_out = d_logits; ga.T = T;
  return launch_gemm_umma(m, ws + wl.h2, P, ws + wl.w2T, Q, ws + wl.dlogits, rows, ga, st);
}

// dlogits -> dp1 -> dskip -> dz planes, and the three bias gradients as column sums
int launch_post_bwd_chain_umma(wn_model* m, unsigned char* ws, int T, float* d_grads, cudaStream_t st) {
  const WorkspaceLayout& wl = m->wl;
  const wn_arch& a = m->a;
  const int64_t rows = (int64_t)m->n_slots * T;
  const int LD = m->L * a.n_dil, S = a.n_skip, P = a.n_post, Q = a.n_quant;
  const bf16* wbf = reinterpret_cast<const bf16*>(ws + wl.wbf);
  int rc;
  {
    ProfScope ps(PROF_POST_BWD, st);
    GemmUmmaArgs ga;
    memset(&ga, 0, sizeof(ga));
    ga.mode = 1;
    ga.H = reinterpret_cast<const bf16*>(ws + wl.h2);
    if ((rc = launch_gemm_umma(m, ws + wl.dlogits, Q, wbf + m->off_post2, P, ws + wl.dp1, rows, ga, st))) return rc;
    ga.H = reinterpret_cast<const bf16*>(ws + wl.h1);
    if ((rc = launch_gemm_umma(m, ws + wl.dp1, P, wbf + m->off_post1, S, ws + wl.dskip, rows, ga, st))) return rc;
    ga.mode = 2;
    ga.H = nullptr;
    ga.D = a.n_dil;
    if ((rc = launch_gemm_umma(m, ws + wl.dskip, S, ws + wl.wsCat, LD, ws + wl.dz, rows, ga, st))) return rc;
    if (a.use_bias) {
      const int ny = (int)std::max<int64_t>(1, std::min<int64_t>(rows / 256, 4 * m->sm_count));
      auto colsum = [&](const void* x, int N, float* out) {
        k_colsum_bf16<<<dim3((N / 2 + 127) / 128, ny), 128, 0, st>>>(reinterpret_cast<const bf16*>(x), rows, N, N, out);
      };
      colsum(ws + wl.dlogits, Q, d_grads + m->off_post2_b);
      WN_LAUNCH_CHECK();
      colsum(ws + wl.dp1, P, d_grads + m->off_post1_b);
      WN_LAUNCH_CHECK();
      colsum(ws + wl.dskip, S, d_grads + m->layers[0].skip_b);
      WN_LAUNCH_CHECK();
    }
  }
  if (a.use_bias && m->L > 1) {
    k_bcast_skip_bias_umma<<<dim3((a.n_skip + 127) / 128, m->L - 1), 128, 0, st>>>(d_grads, m->d_layers, m->L, a.n_skip);
    WN_LAUNCH_CHECK();
  }
  return WN_OK;
}

}  // namespace wn

// =====================================================================================================
// Wide layers (R, D multiples of 64; BASELINE configs[4] has R = D = 128) through k_gemm_umma: forward as two GEMMs
//   v = [x(t-dil) | x(t)] . Wc  -> gate epilogue -> z (stash)          (tmodel.py:117-168)
//   x' = x + z . RESIDUAL + b    -> next layer's input (prefix layout)   (tmodel.py:171-184, :325)
// =====================================================================================================
namespace wn {

// wcP[l][n][k] (K-major B operand, N = 2D rows, K = 2R): row n = h*D + j: j < D/2 -> SIGNAL channel h*D/2 + j,
// else 0.5 * GATE channel h*D/2 + (j - D/2); k < R -> tap 0 (x[t-dil]), else tap 1 (x[t]).  wrT[l][r][d] = RESIDUAL[d][r].
__global__ void k_prep_wide_weights(const float* __restrict__ p, const LayerDesc* __restrict__ layers, int R, int D,
                                    bf16* __restrict__ wcP, bf16* __restrict__ wrT, bf16* __restrict__ wdP) {
  const int l = blockIdx.x;
  const LayerDesc ld = layers[l];
  const int n_wc = 2 * D * 2 * R, n_wr = R * D, Dh = D / 2;
  for (int i = threadIdx.x; i < n_wc; i += blockDim.x) {
    const int n = i / (2 * R), k = i % (2 * R);
    const int h = n / D, j = n % D;
    const bool gate = j >= Dh;
    const int ch = h * Dh + (gate ? j - Dh : j);
    const int tap = k >= R, r = k % R;
    const float v = p[(gate ? ld.gate : ld.sig) + ((int64_t)tap * R + r) * D + ch];
    wcP[(int64_t)l * n_wc + i] = f2bf(gate ? 0.5f * v : v);
  }
  for (int i = threadIdx.x; i < n_wr; i += blockDim.x) {
    const int r = i / D, d = i % D;
    wrT[(int64_t)l * n_wr + i] = f2bf(p[ld.res + (int64_t)d * R + r]);
  }
  // wdP[l][r][k] (N = R rows, K = 4D): k < 2D -> tap 1 (contracted with dv[t]), else tap 0 (with dv[t + dil]);
  // inside a tap: k % 2D < D -> SIGNAL[tap][r][.], else GATE[tap][r][.]
  const int n_wd = R * 4 * D;
  for (int i = threadIdx.x; i < n_wd; i += blockDim.x) {
    const int r = i / (4 * D), k = i % (4 * D);
    const int tap = k < 2 * D ? 1 : 0, kk = k % (2 * D);
    const float v = p[(kk < D ? ld.sig : ld.gate) + ((int64_t)tap * R + r) * D + (kk % D)];
    wdP[(int64_t)l * n_wd + i] = f2bf(v);
  }
}

bool umma_wide_layer_supported(const wn_model* m) {
  static const bool disabled = getenv("WN_DISABLE_UMMA") != nullptr || getenv("WN_DISABLE_UMMA_WIDE") != nullptr;
  const wn_arch& a = m->a;
  return !disabled && a.n_res % 64 == 0 && a.n_dil % 64 == 0 && a.n_res <= 256 && a.n_dil <= 128 && a.n_gc_embed == 0;
}

int launch_prep_wide_umma(wn_model* m, const float* d_params, unsigned char* ws, cudaStream_t st) {
  const WorkspaceLayout& wl = m->wl;
  k_prep_wide_weights<<<m->L, 256, 0, st>>>(d_params, m->d_layers, m->a.n_res, m->a.n_dil,
                                            reinterpret_cast<bf16*>(ws + wl.wcT), reinterpret_cast<bf16*>(ws + wl.wrT),
                                            reinterpret_cast<bf16*>(ws + wl.wdP));
  WN_LAUNCH_CHECK();
  return WN_OK;
}

static int map3(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t slots) {
  const uint64_t dims[3] = {cols, rows, slots};
  const uint64_t strides[2] = {cols * 2, cols * rows * 2};
  const uint32_t box[3] = {UKB, UM, 1};
  return make_tensor_map_bf16(out, base, 3, dims, strides, box, 128);
}

static int launch_gemm_umma_layer(wn_model* m, const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mo,
                                  GemmUmmaArgs ga, int T, cudaStream_t st) {
  ga.rows = (int64_t)m->n_slots * T;
  ga.T = T;
  ga.tps = (T + UM - 1) / UM;
  const size_t smem = gemm_smem_plan(ga);
  WN_CUDA_CHECK(cudaFuncSetAttribute(k_gemm_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  const int n_tiles = ga.tps * m->n_slots;
  int grid = std::max(1, std::min(n_tiles, m->sm_count));
  if (const char* e = getenv("WN_PERSIST_GRID")) if (atoi(e) > 0) grid = std::max(1, std::min(grid, atoi(e)));
  WN_CUDA_CHECK(launch_pdl(k_gemm_umma, grid, UPOST_P_THREADS, smem, st, ma, mb, mo, ga));
  WN_LAUNCH_CHECK();
  return WN_OK;
}

int launch_layer_fwd_wide_umma(wn_model* m, const float* d_params, unsigned char* ws, int T, int l, cudaStream_t st) {
  const WorkspaceLayout& wl = m->wl;
  const wn_arch& a = m->a;
  const LayerDesc& ld = m->layers[l];
  const int R = a.n_res, D = a.n_dil, LD = m->L * D;
  const bool last = (l + 1 == m->L);
  CUtensorMap mx, mz, mwc, mwr, mxo;
  int rc;
  if ((rc = map3(&mx, ws + wl.xfull[l], R, (uint64_t)ld.dil + T, m->n_slots))) return rc;
  if ((rc = map3(&mz, ws + wl.z, LD, T, m->n_slots))) return rc;
  if ((rc = map2d(&mwc, reinterpret_cast<const bf16*>(ws + wl.wcT) + (size_t)l * 2 * D * 2 * R, 2 * R, 2 * D, UKB,
                  (uint32_t)(2 * D)))) return rc;
  ProfScope ps(PROF_LAYER_FWD, st);
  GemmUmmaArgs ga;
  memset(&ga, 0, sizeof(ga));
  ga.mode = 4; ga.K = 2 * R; ga.N = 2 * D;
  ga.a_k_split = R; ga.a_row_off2 = ld.dil;
  ga.out_col0 = l * D; ga.out_row_off = 0;
  ga.bias = ld.sig_b >= 0 ? d_params + ld.sig_b : nullptr;
  ga.bias2 = ld.gate_b >= 0 ? d_params + ld.gate_b : nullptr;
  if ((rc = launch_gemm_umma_layer(m, mx, mwc, mz, ga, T, st))) return rc;
  if (last) return WN_OK;  // the last layer's residual output is unused (tmodel.py:313-325)
  const int dil_next = m->layers[l + 1].dil;
  if ((rc = map2d(&mwr, reinterpret_cast<const bf16*>(ws + wl.wrT) + (size_t)l * R * D, D, R, UKB, (uint32_t)R))) return rc;
  if ((rc = map3(&mxo, ws + wl.xfull[l + 1], R, (uint64_t)dil_next + T, m->n_slots))) return rc;
  memset(&ga, 0, sizeof(ga));
  ga.mode = 5; ga.K = D; ga.N = R;
  ga.a_col0 = l * D;
  ga.out_col0 = 0; ga.out_row_off = dil_next;
  ga.bias = ld.res_b >= 0 ? d_params + ld.res_b : nullptr;
  ga.X = reinterpret_cast<const bf16*>(ws + wl.xfull[l]);
  ga.x_slot_rows = ld.dil + T; ga.x_row_off = ld.dil;
  return launch_gemm_umma_layer(m, mz, mwr, mxo, ga, T, st);
}

}  // namespace wn

namespace wn {

// Backward of one wide layer as three GEMMs (the data gradient in its plain form, as the generation-1 kernels keep it):
//   dz_l      += dx_{l+1} . RESIDUAL^T                      (in place on the layer's dz plane; one extra bf16 rounding)
//   dv         = gate'(v) * dz_l, v recomputed from x        -> dv [B*T][2D]
//   dx_l[t]    = dx_{l+1}[t] + dv[t] . W[1]^T + dv[t + dil] . W[0]^T   (rows t + dil >= T: TMA zero fill = truncated BPTT)
// and the bias gradients as column sums.  The weight gradients follow from k_wgrad_umma (launch_wgrad_umma_x).
int launch_layer_bwd_wide_umma(wn_model* m, const float* d_params, unsigned char* ws, int T, int l, float* d_grads,
                               cudaStream_t st) {
  const WorkspaceLayout& wl = m->wl;
  const wn_arch& a = m->a;
  const LayerDesc& ld = m->layers[l];
  const int R = a.n_res, D = a.n_dil;
  const int64_t rows = (int64_t)m->n_slots * T;
  const bool has_next = (l + 1 < m->L);
  const bf16* wbf = reinterpret_cast<const bf16*>(ws + wl.wbf);
  bf16* dz_plane = reinterpret_cast<bf16*>(ws + wl.dz) + (size_t)l * rows * D;
  bf16* dv = reinterpret_cast<bf16*>(ws + wl.dv);
  bf16* dx_next = has_next ? reinterpret_cast<bf16*>(ws + wl.dx[(l + 1) & 1]) : nullptr;
  bf16* dx_out = reinterpret_cast<bf16*>(ws + wl.dx[l & 1]);
  int rc;
  GemmUmmaArgs ga;
  {
    ProfScope ps(PROF_LAYER_BWD_A, st);
    if (has_next) {  // dz_l += dx_{l+1} . RESIDUAL^T   (RESIDUAL [D][R]: N = D rows, K = R)
      memset(&ga, 0, sizeof(ga));
      ga.mode = 5;
      ga.X = dz_plane;
      if ((rc = launch_gemm_umma(m, dx_next, R, wbf + ld.res, D, dz_plane, rows, ga, st))) return rc;
    }
    CUtensorMap mx, mwc, mdv;
    if ((rc = map3(&mx, ws + wl.xfull[l], R, (uint64_t)ld.dil + T, m->n_slots))) return rc;
    if ((rc = map2d(&mwc, reinterpret_cast<const bf16*>(ws + wl.wcT) + (size_t)l * 2 * D * 2 * R, 2 * R, 2 * D, UKB,
                    (uint32_t)(2 * D)))) return rc;
    if ((rc = map3(&mdv, dv, 2 * D, T, m->n_slots))) return rc;
    memset(&ga, 0, sizeof(ga));
    ga.mode = 6; ga.K = 2 * R; ga.N = 2 * D;
    ga.a_k_split = R; ga.a_row_off2 = ld.dil;
    ga.bias = ld.sig_b >= 0 ? d_params + ld.sig_b : nullptr;
    ga.bias2 = ld.gate_b >= 0 ? d_params + ld.gate_b : nullptr;
    ga.X = dz_plane;
    if (a.use_bias) {  // SIGNAL_BIAS / GATE_BIAS gradients = column sums of dv, taken from the staged output tiles
      ga.colsum_out = d_grads + ld.sig_b;
      ga.colsum_out2 = d_grads + ld.gate_b;
      ga.colsum_split = D;
    }
    if ((rc = launch_gemm_umma_layer(m, mx, mwc, mdv, ga, T, st))) return rc;
  }
  {
    ProfScope ps(PROF_LAYER_BWD_B, st);
    CUtensorMap mdv, mwd, mdx;
    if ((rc = map3(&mdv, dv, 2 * D, T, m->n_slots))) return rc;
    if ((rc = map2d(&mwd, reinterpret_cast<const bf16*>(ws + wl.wdP) + (size_t)l * R * 4 * D, 4 * D, R, UKB, (uint32_t)R)))
      return rc;
    if ((rc = map3(&mdx, dx_out, R, T, m->n_slots))) return rc;
    memset(&ga, 0, sizeof(ga));
    ga.mode = 5; ga.K = 4 * D; ga.N = R;
    ga.a_k_split = 2 * D; ga.a_row_off2 = ld.dil;
    ga.X = dx_next;
    ga.x_slot_rows = T; ga.x_row_off = 0;
    if (a.use_bias && l > 0) {  // RESIDUAL_BIAS gradient of the layer below = column sums of the dx_l produced here
      ga.colsum_out = d_grads + m->layers[l - 1].res_b;
      ga.colsum_out2 = ga.colsum_out;
      ga.colsum_split = R;
    }
    if ((rc = launch_gemm_umma_layer(m, mdv, mwd, mdx, ga, T, st))) return rc;
  }
  return WN_OK;
}

}  // namespace wn

// =====================================================================================================
// k_wgrad_umma: dW[m][n] += sum_rows A[row][a_col0 + m] * Y[row][n]      (weight gradients, split-K)
// Both operands are MN-major for the tensor core (the contraction index -- time -- is the slow one in
// memory), staged by TMA as 64-column SW128 panels.  One CTA owns a 128 x N output tile (N <= 256) in TMEM
// and a contiguous range of 64-row K blocks; partial sums are reduced with coalesced fp32 atomics.
// =====================================================================================================
namespace wn {

constexpr int WG_BK = 64;                 // rows (timesteps) per K block
constexpr int WG_PANEL = WG_BK * 128;     // one 64-column panel: 8 KB
constexpr int WG_STAGES = 4;
constexpr int WG_A_BYTES = 2 * WG_PANEL;  // M = 128
constexpr int WG_B_BYTES = 4 * WG_PANEL;  // N <= 256

struct WgradUmmaArgs {
  float* out;                 // mode 0: out[m * ldo + n]
  const LayerDesc* layers;    // mode 1: row m = l*D + d -> grads + layers[l].skip + d * ldo
  float* grads;
  int mode, D, ldo;
  int M_total, N;
  int a_col0;
  int out_col0;  // mode 1: first output column (N chunks of a wider gradient)
  int n_taps, m_tiles;  // n_taps == 2: blockIdx.x / m_tiles is the conv tap: A rows shifted by tap * a_tap_row_off,
  int a_tap_row_off;    // outputs by tap * tap_out_stride (both taps of a layer share Y = dv: one launch, half the splits)
  int64_t tap_out_stride;
  float* out2;         // mode 0 with n_split > 0: columns >= n_split go to out2[m * ldo + (n - n_split)]
  int n_split;
  int a_T, a_row_off;  // a_T > 0: A is a layer input in the prefix layout [slot][dil + T][R] (3-D map): flat row r is
                       // (slot r / a_T, row r % a_T + a_row_off); a_T % 64 == 0 keeps a K block inside one slot
  int64_t kblocks_total, kblocks_per_cta;
  long long* trace;  // wn_debug_trace(buf, -3) (tools/trace_layer.py wgrad): timeline of CTA 0 of the LAST launch
};

// 16-byte vector reduction into global memory (sm_90+): one L2 operation for four fp32 adds
__device__ __forceinline__ void red_add_v4(float* dst, float v0, float v1, float v2, float v3) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v0), "f"(v1), "f"(v2), "f"(v3) : "memory");
}

__global__ void __launch_bounds__(UPOST_THREADS, 1)
k_wgrad_umma(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_y, WgradUmmaArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* stage_a = smem;
  unsigned char* stage_b = smem + WG_STAGES * WG_A_BYTES;
  __shared__ __align__(8) uint64_t full_bar[WG_STAGES], empty_bar[WG_STAGES], acc_full;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tap = a.n_taps == 2 ? (int)blockIdx.x / a.m_tiles : 0;
  const int m0 = (a.n_taps == 2 ? (int)blockIdx.x % a.m_tiles : (int)blockIdx.x) * 128;
  const int a_row_off = a.a_row_off + tap * a.a_tap_row_off;
  float* const out = a.out + tap * a.tap_out_stride;
  float* const out2 = a.out2 != nullptr ? a.out2 + tap * a.tap_out_stride : nullptr;
  const int64_t kb0 = (int64_t)blockIdx.y * a.kblocks_per_cta;
  const int64_t kb1 = min(a.kblocks_total, kb0 + a.kblocks_per_cta);
  const int nkb = (int)max((int64_t)0, kb1 - kb0);
  const int N = a.N, npan = (N + 63) / 64;
  uint32_t ncols = 32;
  while (ncols < (uint32_t)N) ncols <<= 1;

  pdl_launch_dependents();
  PostTracer tr;
  tr.init(a.trace, warp, blockIdx.x == 0 && blockIdx.y == 0 && lane == 0);
  tr.ev(30, 0);
  if (tid == 0) {
    for (int i = 0; i < WG_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    mbar_init(&acc_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_y);
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, ncols);
  pdl_wait();
  tr.ev(31, 0);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;

  if (nkb > 0) {
    if (warp == 0) {
      if (lane == 0) {
        for (int it = 0; it < nkb; ++it) {
          const int st = it % WG_STAGES;
          if ((it & 15) == 0) tr.ev(1, it);
          mbar_wait(&empty_bar[st], ((uint32_t)(it / WG_STAGES) & 1u) ^ 1u);
          if ((it & 15) == 0) tr.ev(2, it);
          mbar_expect_tx(&full_bar[st], (uint32_t)((2 + npan) * WG_PANEL));
          const int row = (int)((kb0 + it) * WG_BK);
          for (int c = 0; c < 2; ++c) {
            if (a.a_T > 0)
              tma_load_3d(stage_a + st * WG_A_BYTES + c * WG_PANEL, &map_a, &full_bar[st], a.a_col0 + m0 + c * 64,
                          row % a.a_T + a_row_off, row / a.a_T);
            else
              tma_load_2d(stage_a + st * WG_A_BYTES + c * WG_PANEL, &map_a, &full_bar[st], a.a_col0 + m0 + c * 64, row);
          }
          for (int c = 0; c < npan; ++c)
            tma_load_2d(stage_b + st * WG_B_BYTES + c * WG_PANEL, &map_y, &full_bar[st], c * 64, row);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N, true, true);
        for (int it = 0; it < nkb; ++it) {
          const int st = it % WG_STAGES;
          if ((it & 15) == 0) tr.ev(3, it);
          mbar_wait(&full_bar[st], (uint32_t)(it / WG_STAGES) & 1u);
          if ((it & 15) == 0) tr.ev(16, it);
          tc_fence_after_sync();
          const uint32_t sa = smem_u32(stage_a + st * WG_A_BYTES), sb = smem_u32(stage_b + st * WG_B_BYTES);
#pragma unroll
          for (int k = 0; k < WG_BK / 16; ++k)
            mma_bf16_ss(tmem_base, make_mnmajor_desc(sa + k * 16 * 128, 128, WG_PANEL),
                        make_mnmajor_desc(sb + k * 16 * 128, 128, WG_PANEL), idesc, (it | k) != 0);
          mma_commit(&empty_bar[st]);
        }
        mma_commit(&acc_full);
      }
    } else {
      // epilogue: TMEM -> fp32 staging in the (now idle) pipeline buffers -> coalesced atomics
      const int q4 = warp & 3;
      const int r = q4 * 32 + lane;
      const int et = (warp - 2) * 32 + lane;
      float* stg = reinterpret_cast<float*>(smem);  // [128][N + 1]
      const int lds = N + 1;
      tr.ev(5, 0);
      mbar_wait(&acc_full, 0);
      tr.ev(6, 0);
      tc_fence_after_sync();
      uint32_t v[32];
      for (int c0 = 0; c0 < N; c0 += 32) {
        tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j < N) stg[r * lds + c0 + j] = __uint_as_float(v[j]);
      }
      epi_bar_sync();
      // four consecutive columns per thread: one 16-byte vector reduction when the destination allows it (a quarter of
      // the L2 atomic operations: with 148 splits per output tile the kernel is bound by them, not by the GEMM)
      for (int idx = et * 4; idx < 128 * N; idx += 128 * 4) {
        const int rr = idx / N, c = idx % N;  // N % 16 == 0: the four columns stay inside the row
        const int m = m0 + rr;
        if (m < a.M_total) {
          const float* sp = stg + rr * lds + c;
          const float v0 = sp[0], v1 = sp[1], v2 = sp[2], v3 = sp[3];
          const bool second = a.mode == 0 && a.n_split > 0 && c >= a.n_split;
          float* dst = a.mode == 0 ? (second ? out2 + (size_t)m * a.ldo + (c - a.n_split) : out + (size_t)m * a.ldo + c)
                                   : a.grads + a.layers[m / a.D].skip + (size_t)(m % a.D) * a.ldo + a.out_col0 + c;
          const bool same_dst = a.mode != 0 || a.n_split <= 0 || second || c + 3 < a.n_split;
          if (same_dst && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            if (v0 != 0.f || v1 != 0.f || v2 != 0.f || v3 != 0.f) red_add_v4(dst, v0, v1, v2, v3);
          } else {
            const float vv[4] = {v0, v1, v2, v3};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int cj = c + j;
              float* dj = a.mode == 0 && a.n_split > 0 && cj >= a.n_split ? out2 + (size_t)m * a.ldo + (cj - a.n_split)
                          : a.mode == 0                                   ? out + (size_t)m * a.ldo + cj
                                                                          : dst + j;
              if (vv[j] != 0.f) atomicAdd(dj, vv[j]);
            }
          }
        }
      }
    }
  }
  tr.ev(32, 0);
  tc_fence_before_sync();
  __syncthreads();
  tr.ev(33, 0);
  if (warp == 1) tmem_dealloc(tmem_base, ncols);
}

// A: [rows][lda] bf16 (columns a_col0 .. a_col0 + M_total), Y: [rows][N] bf16 (row pitch ldy)
static int launch_wgrad_umma_at(wn_model* m, const bf16* A, int lda, int a_col0, int M_total, const bf16* Y, int ldy,
                                int N, int64_t rows, float* out, int ldo, int mode, float* grads, int out_col0,
                                cudaStream_t st);

int launch_wgrad_umma(wn_model* m, const bf16* A, int lda, int a_col0, int M_total, const bf16* Y, int ldy, int N,
                      int64_t rows, float* out, int ldo, int mode, float* grads, cudaStream_t st) {
  return launch_wgrad_umma_at(m, A, lda, a_col0, M_total, Y, ldy, N, rows, out, ldo, mode, grads, 0, st);
}

// the same for N > 256: one launch per chunk of 256 output columns (Y and the output advance by the chunk)
int launch_wgrad_umma_cols(wn_model* m, const bf16* A, int lda, int M_total, const bf16* Y, int N_total, int64_t rows,
                           float* out, int mode, float* grads, cudaStream_t st) {
  for (int n0 = 0; n0 < N_total; n0 += 256) {
    const int rc = launch_wgrad_umma_at(m, A, lda, 0, M_total, Y + n0, N_total, std::min(256, N_total - n0), rows,
                                        out != nullptr ? out + n0 : nullptr, N_total, mode, grads, n0, st);
    if (rc) return rc;
  }
  return WN_OK;
}

static int launch_wgrad_umma_at(wn_model* m, const bf16* A, int lda, int a_col0, int M_total, const bf16* Y, int ldy,
                                int N, int64_t rows, float* out, int ldo, int mode, float* grads, int out_col0,
                                cudaStream_t st) {
  CUtensorMap ma, my;
  int rc;
  if ((rc = map2d(&ma, A, (uint64_t)lda, (uint64_t)rows, 64, WG_BK))) return rc;
  if ((rc = map2d(&my, Y, (uint64_t)ldy, (uint64_t)rows, 64, WG_BK))) return rc;
  WgradUmmaArgs wa;
  memset(&wa, 0, sizeof(wa));
  wa.out = out; wa.layers = m->d_layers; wa.grads = grads; wa.mode = mode; wa.D = m->a.n_dil; wa.ldo = ldo;
  wa.M_total = M_total; wa.N = N; wa.a_col0 = a_col0; wa.out_col0 = out_col0;
  wa.kblocks_total = (rows + WG_BK - 1) / WG_BK;
  wa.trace = g_trace_layer == -3 ? g_trace_buf : nullptr;
  const int m_tiles = (M_total + 127) / 128;
  const int sms = std::max(1, m->sm_count);
  int64_t splits = std::max<int64_t>(1, std::min<int64_t>(wa.kblocks_total, sms / m_tiles));
  wa.kblocks_per_cta = (wa.kblocks_total + splits - 1) / splits;
  splits = (wa.kblocks_total + wa.kblocks_per_cta - 1) / wa.kblocks_per_cta;
  const size_t smem = (size_t)WG_STAGES * (WG_A_BYTES + WG_B_BYTES) + 1024;
  WN_CUDA_CHECK(cudaFuncSetAttribute(k_wgrad_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps(PROF_WGRAD, st);
  WN_CUDA_CHECK(launch_pdl(k_wgrad_umma, dim3(m_tiles, (unsigned)splits), UPOST_THREADS, smem, st, ma, my, wa));
  WN_LAUNCH_CHECK();
  return WN_OK;
}

// Weight gradients of one conv tap on the layers the fused tcgen05 backward does not cover (R, D multiples of 64):
// dW[r][n] = sum_t x_tap[t][r] * dv[t][y_col0 + n], x read straight from the prefix layout xfull_l; tap < 0: both taps
// in one launch (outputs of tap 1 at out + tap_out_stride)
bool umma_wgrad_x_supported(const wn_model* m, int T) {
  static const bool disabled = getenv("WN_DISABLE_UMMA") != nullptr;
  const wn_arch& a = m->a;
  return !disabled && a.n_res % 64 == 0 && a.n_dil % 64 == 0 && a.n_res <= 256 && a.n_dil <= 256 && T % 64 == 0;
}

int launch_wgrad_umma_x(wn_model* m, const bf16* xfull, int dil, int T, int tap, const bf16* Y, int ldy, int N,
                        float* out, float* out2, int n_split, int ldo, int64_t tap_out_stride, cudaStream_t st) {
  const int R = m->a.n_res;
  const int64_t rows = (int64_t)m->n_slots * T;
  CUtensorMap ma, my;
  int rc;
  {
    const uint64_t dims[3] = {(uint64_t)R, (uint64_t)(dil + T), (uint64_t)m->n_slots};
    const uint64_t strides[2] = {(uint64_t)R * 2, (uint64_t)(dil + T) * R * 2};
    const uint32_t box[3] = {64, WG_BK, 1};
    if ((rc = make_tensor_map_bf16(&ma, xfull, 3, dims, strides, box, 128))) return rc;
  }
  if ((rc = map2d(&my, Y, (uint64_t)ldy, (uint64_t)rows, 64, WG_BK))) return rc;
  WgradUmmaArgs wa;
  memset(&wa, 0, sizeof(wa));
  wa.out = out; wa.layers = m->d_layers; wa.mode = 0; wa.D = m->a.n_dil; wa.ldo = ldo;
  wa.M_total = R; wa.N = N; wa.a_T = T; wa.a_row_off = tap > 0 ? dil : 0;
  wa.out2 = out2; wa.n_split = n_split;
  wa.kblocks_total = (rows + WG_BK - 1) / WG_BK;
  const int n_taps = tap < 0 ? 2 : 1;
  const int m_tiles = (R + 127) / 128 * n_taps;
  wa.n_taps = n_taps; wa.m_tiles = m_tiles / n_taps; wa.a_tap_row_off = dil; wa.tap_out_stride = tap_out_stride;
  const int sms = std::max(1, m->sm_count);
  int64_t splits = std::max<int64_t>(1, std::min<int64_t>(wa.kblocks_total, sms / m_tiles));
  wa.kblocks_per_cta = (wa.kblocks_total + splits - 1) / splits;
  splits = (wa.kblocks_total + wa.kblocks_per_cta - 1) / wa.kblocks_per_cta;
  const size_t smem = (size_t)WG_STAGES * (WG_A_BYTES + WG_B_BYTES) + 1024;
  WN_CUDA_CHECK(cudaFuncSetAttribute(k_wgrad_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps(PROF_WGRAD, st);
  WN_CUDA_CHECK(launch_pdl(k_wgrad_umma, dim3(m_tiles, (unsigned)splits), UPOST_THREADS, smem, st, ma, my, wa));
  WN_LAUNCH_CHECK();
  return WN_OK;
}

bool umma_wgrad_supported(const wn_model* m, int lda, int ldy, int N) {
  return umma_post_supported(m) && N % 16 == 0 && N <= 256 && lda % 8 == 0 && ldy % 8 == 0;
}

// =====================================================================================================
// Local conditioning (reference tmodel.py:68-83 _preprocess_lc, :156-160 the per-layer LC projections; arch.py:75-80).
// A transposed convolution whose width equals its stride is a plain GEMM: level i maps [B*T_i x n_in] onto
// [B*T_i x s_i*n_out], which IS [B*T_i*s_i x n_out] row-major (out[b, t*s + k, o] = sum_c in[b, t, c] W_i[k][o][c]).
// All LC activations use 128-column rows (LCP; channels beyond n_lc_in / n_lc_out are zero), so every level is
// k_gemm_umma with K = 128, N = s_i * 128; the per-layer projections of ALL layers are one more GEMM
// lc_up[rows x 128] . [LC_SIGNAL_l | LC_GATE_l]_l (N = L * 2D) whose epilogue writes per-layer planes cond[l][rows][2D]
// -- the operand the fused layer kernels add in their gate epilogues.  Backward: the layer kernels leave dv in those
// planes; LC_SIGNAL / LC_GATE gradients = lc_up^T . dcond_l (split-K k_wgrad_umma), d lc_up = sum_l dcond_l . Wlc_l^T
// (one GEMM whose K blocks are the planes), then the upsampling chain in reverse.
// =====================================================================================================
constexpr int LCP = 128;

struct LcPrepArgs {
  int n_levels, L, D, n_in, n_out;
  int s[8];
  int64_t off_up[8];
  bf16* wup[8];
  bf16* wupT[8];
  bf16* wcat;
  bf16* wcatT;
};

// blocks [0, n_levels): level i ; blocks [n_levels, n_levels + L): layer l
__global__ void k_lc_prep(const float* __restrict__ p, const LayerDesc* __restrict__ layers, LcPrepArgs a) {
  const int bid = blockIdx.x;
  if (bid < a.n_levels) {
    const int i = bid, s = a.s[i], nin = i == 0 ? a.n_in : a.n_out;
    const float* w = p + a.off_up[i];  // [s][n_out][nin]
    for (int e = threadIdx.x; e < s * LCP * LCP; e += blockDim.x) {
      const int n = e / LCP, c = e % LCP;  // n = k * 128 + o
      const int k = n / LCP, o = n % LCP;
      const float v = (o < a.n_out && c < nin) ? w[((int64_t)k * a.n_out + o) * nin + c] : 0.f;
      a.wup[i][e] = f2bf(v);
      a.wupT[i][(int64_t)c * (s * LCP) + n] = f2bf(v);
    }
  } else {
    const int l = bid - a.n_levels, D = a.D;
    const LayerDesc ld = layers[l];
    const int64_t LD2 = (int64_t)a.L * 2 * D;
    for (int e = threadIdx.x; e < 2 * D * LCP; e += blockDim.x) {
      const int n = e / LCP, c = e % LCP;
      const float v = c < a.n_out ? p[(n < D ? ld.lc_sig : ld.lc_gate) + (int64_t)c * D + (n % D)] : 0.f;
      a.wcat[((int64_t)l * 2 * D + n) * LCP + c] = f2bf(v);
      a.wcatT[(int64_t)c * LD2 + (int64_t)l * 2 * D + n] = f2bf(v);
    }
  }
}

__global__ void k_lc_mel(const float* __restrict__ mel, bf16* __restrict__ out, int64_t rows0, int n_in) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows0 * LCP) return;
  const int64_t r = i / LCP;
  const int c = (int)(i % LCP);
  out[i] = f2bf(c < n_in ? mel[r * n_in + c] : 0.f);
}

// grads <- the fp32 scratch the split-K kernels accumulated into
__global__ void k_lc_scatter(const float* __restrict__ gtmp, const LayerDesc* __restrict__ layers, float* __restrict__ grads,
                             LcPrepArgs a, int64_t up_off0) {
  const int bid = blockIdx.x;
  if (bid < a.L) {
    const LayerDesc ld = layers[bid];
    const float* g = gtmp + (int64_t)bid * LCP * 2 * a.D;  // [128][2D]
    for (int e = threadIdx.x; e < a.n_out * 2 * a.D; e += blockDim.x) {
      const int c = e / (2 * a.D), n = e % (2 * a.D);
      grads[(n < a.D ? ld.lc_sig : ld.lc_gate) + (int64_t)c * a.D + (n % a.D)] = g[(int64_t)c * 2 * a.D + n];
    }
  } else {
    const int i = bid - a.L, s = a.s[i], nin = i == 0 ? a.n_in : a.n_out;
    int64_t off = up_off0;
    for (int j = 0; j < i; ++j) off += (int64_t)LCP * a.s[j] * LCP;
    const float* g = gtmp + off;  // [128 (c)][s * 128 (k * 128 + o)]
    for (int e = threadIdx.x; e < s * a.n_out * nin; e += blockDim.x) {
      const int k = e / (a.n_out * nin), o = (e / nin) % a.n_out, c = e % nin;
      grads[a.off_up[i] + e] = g[(int64_t)c * (s * LCP) + k * LCP + o];
    }
  }
}

static LcPrepArgs lc_args(wn_model* m, unsigned char* ws) {
  const WorkspaceLayout& wl = m->wl;
  LcPrepArgs a;
  memset(&a, 0, sizeof(a));
  a.n_levels = m->a.n_lc_layers; a.L = m->L; a.D = m->a.n_dil; a.n_in = m->a.n_lc_in; a.n_out = m->a.n_lc_out;
  for (int i = 0; i < a.n_levels; ++i) {
    a.s[i] = m->a.lc_upsample[i];
    a.off_up[i] = m->off_lc_up[i];
    a.wup[i] = reinterpret_cast<bf16*>(ws + wl.lc_wup[i]);
    a.wupT[i] = reinterpret_cast<bf16*>(ws + wl.lc_wupT[i]);
  }
  a.wcat = reinterpret_cast<bf16*>(ws + wl.lc_wcat);
  a.wcatT = reinterpret_cast<bf16*>(ws + wl.lc_wcatT);
  return a;
}

// mel frames -> upsampled conditioning -> per-layer planes cond[l][rows][2D]  (runs before the layer loop)
int launch_lc_fwd(wn_model* m, const float* d_params, const float* d_mel, unsigned char* ws, int T, cudaStream_t st) {
  const WorkspaceLayout& wl = m->wl;
  const wn_arch& a = m->a;
  const int64_t rows = (int64_t)m->n_slots * T;
  const int n = a.n_lc_layers;
  int rc;
  ProfScope ps(PROF_PREP, st);
  const LcPrepArgs pa = lc_args(m, ws);
  k_lc_prep<<<n + m->L, 256, 0, st>>>(d_params, m->d_layers, pa);
  WN_LAUNCH_CHECK();
  int64_t r = rows / m->lc_hop;
  k_lc_mel<<<(unsigned)((r * LCP + 255) / 256), 256, 0, st>>>(d_mel, reinterpret_cast<bf16*>(ws + wl.lc_x[0]), r, a.n_lc_in);
  WN_LAUNCH_CHECK();
  GemmUmmaArgs ga;
  for (int i = 0; i < n; ++i) {
    memset(&ga, 0, sizeof(ga));
    ga.mode = 7;
    if ((rc = launch_gemm_umma(m, ws + wl.lc_x[i], LCP, ws + wl.lc_wup[i], a.lc_upsample[i] * LCP, ws + wl.lc_x[i + 1], r, ga, st)))
      return rc;
    r *= a.lc_upsample[i];
  }
  memset(&ga, 0, sizeof(ga));
  ga.mode = 2;
  ga.D = 2 * a.n_dil;
  return launch_gemm_umma(m, ws + wl.lc_x[n], LCP, ws + wl.lc_wcat, m->L * 2 * a.n_dil, ws + wl.cond, rows, ga, st);
}

// after every layer's backward has left dv in its plane: LC_SIGNAL / LC_GATE / LC_UPSAMPLE gradients
int launch_lc_bwd(wn_model* m, unsigned char* ws, int T, float* d_grads, cudaStream_t st) {
  const WorkspaceLayout& wl = m->wl;
  const wn_arch& a = m->a;
  const int64_t rows = (int64_t)m->n_slots * T;
  const int n = a.n_lc_layers, D2 = 2 * a.n_dil;
  int rc;
  int64_t gt = (int64_t)m->L * LCP * D2;
  const int64_t up_off0 = gt;
  for (int i = 0; i < n; ++i) gt += (int64_t)LCP * a.lc_upsample[i] * LCP;
  float* gtmp = reinterpret_cast<float*>(ws + wl.lc_gtmp);
  WN_CUDA_CHECK(cudaMemsetAsync(gtmp, 0, sizeof(float) * gt, st));
  const bf16* lc_up = reinterpret_cast<const bf16*>(ws + wl.lc_x[n]);
  const bf16* dcond = reinterpret_cast<const bf16*>(ws + wl.dcond);
  for (int l = 0; l < m->L; ++l)
    if ((rc = launch_wgrad_umma(m, lc_up, LCP, 0, LCP, dcond + (size_t)l * rows * D2, D2, D2, rows,
                                gtmp + (size_t)l * LCP * D2, D2, 0, nullptr, st)))
      return rc;
  GemmUmmaArgs ga;
  memset(&ga, 0, sizeof(ga));
  ga.mode = 7;
  ga.a_planes = 1;
  if ((rc = launch_gemm_umma(m, ws + wl.dcond, m->L * D2, ws + wl.lc_wcatT, LCP, ws + wl.lc_dx[n], rows, ga, st))) return rc;
  std::vector<int64_t> rws(n + 1);
  rws[0] = rows / m->lc_hop;
  for (int i = 0; i < n; ++i) rws[i + 1] = rws[i] * a.lc_upsample[i];
  int64_t up_off = up_off0;
  std::vector<int64_t> up_offs(n);
  for (int i = 0; i < n; ++i) { up_offs[i] = up_off; up_off += (int64_t)LCP * a.lc_upsample[i] * LCP; }
  for (int i = n - 1; i >= 0; --i) {
    const int N = a.lc_upsample[i] * LCP;
    if ((rc = launch_wgrad_umma_cols(m, reinterpret_cast<const bf16*>(ws + wl.lc_x[i]), LCP, LCP,
                                     reinterpret_cast<const bf16*>(ws + wl.lc_dx[i + 1]), N, rws[i], gtmp + up_offs[i], 0,
                                     nullptr, st)))
      return rc;
    if (i > 0) {
      memset(&ga, 0, sizeof(ga));
      ga.mode = 7;
      if ((rc = launch_gemm_umma(m, ws + wl.lc_dx[i + 1], N, ws + wl.lc_wupT[i], LCP, ws + wl.lc_dx[i], rws[i], ga, st))) return rc;
    }
  }
  const LcPrepArgs pa = lc_args(m, ws);
  k_lc_scatter<<<m->L + n, 256, 0, st>>>(gtmp, m->d_layers, d_grads, pa, up_off0);
  WN_LAUNCH_CHECK();
  return WN_OK;
}

}  // namespace wn

// C[M][N] (fp32, accumulated into: zero it first) += A[K][M]^T . B[K][N]; both operands MN-major
extern "C" int wn_selftest_umma_gemm_tn(const void* d_a, const void* d_b, float* d_c, int32_t M, int32_t N,
                                        int32_t K, void* stream_) {
  using namespace wn;
  if (!d_a || !d_b || !d_c || M < 1 || N < 16 || N > 256 || N % 16 || M % 8 || K < 1) {
    set_error("wn_selftest_umma_gemm_tn: invalid argument");
    return WN_ERR_INVALID;
  }
  wn_model fake;
  fake.a.n_dil = 1;
  fake.d_layers = nullptr;
  int dev = 0;
  WN_CUDA_CHECK(cudaGetDevice(&dev));
  WN_CUDA_CHECK(cudaDeviceGetAttribute(&fake.sm_count, cudaDevAttrMultiProcessorCount, dev));
  return launch_wgrad_umma(&fake, reinterpret_cast<const bf16*>(d_a), M, 0, M, reinterpret_cast<const bf16*>(d_b), N, N,
                           K, d_c, N, 0, nullptr, (cudaStream_t)stream_);
}
