// Per-layer kernels, tcgen05 generation (R = D = 32: one activation row is one 64-byte swizzle span).
//
// Activations live in the "prefix" layout xfull_l [slot][dil_l + T][R] (rows [0, dil) = saved D-separation
// state, reference tmodel.py:127), so both conv taps are plain TMA boxes of the same 3-D tensor at row t0
// (x[t-dil]) and t0 + dil (x[t]).
//
// k_layer_fwd_p_umma  (reference tmodel.py:117-168 _dilated_conv, :171-184 _chan_reduce, :325 residual add)
//     acc_v[128 x 2D] = x[t-dil] . W[0] + x[t] . W[1] + 1 . bias      (SIGNAL | GATE side by side in N)
//     z = bf16(tanh(v_s) * sigmoid(v_g))                               -> smem (A of the next MMA) + TMA store
//     acc_r[128 x R]  = z . RESIDUAL + x[t] . I + 1 . bias             -> x' = bf16(acc_r), TMA store into xfull_{l+1}
//   Everything that used to be epilogue arithmetic on CUDA cores (bias adds, the residual add, the 0.5 scaling
//   inside sigmoid(x) = 0.5 tanh(0.5 x) + 0.5) is folded into the tensor-core contraction: biases enter as one
//   extra K = 16 step against a constant-one A tile (hi + lo bf16 split, error 2^-17 relative), the residual as
//   x[t] times a bf16 identity (exact), and the GATE filter/bias copies are pre-scaled by 0.5 (exact).
//
// k_layer_bwd_fused_umma: the whole backward of one layer in ONE persistent kernel.  The data gradient is kept
//   in split form  dx_l[t] = Y_l[t] + P0_l[t + dil_l]  so that no tile ever needs another tile's result:
//     acc_v = x[t-dil] . W[0] + x[t] . W[1] + bias                    (recomputed pre-activations)
//     acc_d = Y_{l+1}[t] . RESIDUAL^T + P0_{l+1}[t+dil_{l+1}] . RESIDUAL^T   (residual part of dz; linearity: the merged
//                                                                      dx_{l+1} is never materialised)
//     th, sg, z ;  dz = dz_skip + acc_d ;  dv = [dz sg (1-th^2) | dz th sg (1-sg)]        (bf16 tile, smem ONLY)
//     acc_p = dv . [W[0]^T | W[1]^T]
//     P0_l = bf16(acc_p[:, :R]) ;  Y_l = bf16(acc_p[:, R:] + Y_{l+1} + P0_{l+1}[t+dil])   -> TMA stores (the only HBM writes)
//   weight gradients accumulate in tensor memory across ALL tiles of the CTA (the K-major activation tiles are
//   re-described MN-major, no extra traffic):  acc_w += [x[t-dil] | x[t] | Y_{l+1} | P0_{l+1}]^T . [dv | z]
//   (rows 64..95 + rows 96..127 = RESIDUAL's gradient, transposed);  acc_b += [dv | ..]^T . 1  = the bias gradients.
//   Rows t + dil >= T of P0 are TMA zero fill: the gradient stops at the stage boundary (SAVE is a variable, not
//   a graph tensor, reference tmodel.py:123-124,165).
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "umma.cuh"

namespace wn {

using namespace umma;

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// write NB bytes (multiple of 16) of one row into a K-major swizzled tile whose rows are SPAN bytes
template <int SPAN, int NB>
__device__ __forceinline__ void row_store(unsigned char* tile, int row, int byte0, const uint32_t* pk) {
#pragma unroll
  for (int ch = 0; ch < NB / 16; ++ch) {
    const uint32_t off = swizzled_offset((uint32_t)row, (uint32_t)(byte0 + ch * 16), SPAN);
    *reinterpret_cast<uint4*>(tile + off) = make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
  }
}
template <int SPAN, int NB>
__device__ __forceinline__ void row_load(const unsigned char* tile, int row, int byte0, uint32_t* pk) {
#pragma unroll
  for (int ch = 0; ch < NB / 16; ++ch) {
    const uint32_t off = swizzled_offset((uint32_t)row, (uint32_t)(byte0 + ch * 16), SPAN);
    const uint4 v = *reinterpret_cast<const uint4*>(tile + off);
    pk[ch * 4] = v.x; pk[ch * 4 + 1] = v.y; pk[ch * 4 + 2] = v.z; pk[ch * 4 + 3] = v.w;
  }
}

// ---- small operand tiles the CTA builds for itself ----------------------------------------------------
// B operand [n_rows][16] bf16, K-major, 32-byte rows (SW32): column 0 = bf16(b), column 1 = bf16(b - hi), rest 0.
// Contracted against a constant-one A tile it adds b (to 2^-17 relative) to every row of the accumulator.
template <typename F>
__device__ __forceinline__ void build_bias_tile(unsigned char* tile, int n_rows, int tid, int nthreads, F bias_of) {
  for (int n = tid; n < n_rows; n += nthreads) {
    const float b = bias_of(n);
    const bf16 hi = f2bf(b);
    const bf16 lo = f2bf(b - bf2f(hi));
    const uint32_t w = (uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16);
    *reinterpret_cast<uint4*>(tile + swizzled_offset((uint32_t)n, 0, 32)) = make_uint4(w, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(tile + swizzled_offset((uint32_t)n, 16, 32)) = make_uint4(0u, 0u, 0u, 0u);
  }
}
// bf16 identity [32][32], K-major, 64-byte rows (SW64)
__device__ __forceinline__ void build_identity_tile(unsigned char* tile, int tid) {
  if (tid < 32) {
    const int n = tid;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      if (ch == (n >> 3)) w[(n & 7) >> 1] = 0x3f80u << (16 * (n & 1));
      *reinterpret_cast<uint4*>(tile + swizzled_offset((uint32_t)n, (uint32_t)(ch * 16), 64)) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}
// constant 1.0 (bf16 0x3f80): the swizzle permutes equal values, so a plain fill is exact
__device__ __forceinline__ void fill_ones(unsigned char* tile, int bytes, int tid, int nthreads) {
  for (int i = tid; i < bytes / 16; i += nthreads)
    *reinterpret_cast<uint4*>(tile + i * 16) = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
}

// ---- weight preparation: K-major B operands -------------------------------------------------------
// wcT[l][tap][n][r] = (n < D ? SIGNAL : 0.5 * GATE)[tap][r][n % D]  ([2D rows][R], one block per tap)
// wrT[l][r][d]      = RESIDUAL[d][r]                               ([R rows][D])
// wr [l][d][r]      = RESIDUAL[d][r] (bf16 copy, [D rows][R]: B operand of dz = dx' . RESIDUAL^T)
// The 0.5 on the GATE half implements sigmoid(g) = 0.5 tanh(0.5 g) + 0.5 without a multiply in the epilogue
// (bf16(0.5 w) == 0.5 bf16(w)).
__global__ void k_prep_layer_weights(const float* __restrict__ p, const LayerDesc* __restrict__ layers, int L, int R,
                                     int D, bf16* __restrict__ wcT, bf16* __restrict__ wrT, bf16* __restrict__ wr) {
  const int l = blockIdx.x;
  const LayerDesc ld = layers[l];
  const int n_wc = 2 * 2 * D * R, n_wr = R * D;
  for (int i = threadIdx.x; i < n_wc; i += blockDim.x) {
    const int tap = i / (2 * D * R), rem = i % (2 * D * R);
    const int n = rem / R, r = rem % R;
    const float v = p[(n < D ? ld.sig : ld.gate) + ((int64_t)tap * R + r) * D + (n % D)];
    wcT[(int64_t)l * n_wc + i] = f2bf(n < D ? v : 0.5f * v);
  }
  for (int i = threadIdx.x; i < n_wr; i += blockDim.x) {
    const int r = i / D, d = i % D;
    const float v = p[ld.res + (int64_t)d * R + r];
    wrT[(int64_t)l * n_wr + i] = f2bf(v);
    wr[(int64_t)l * n_wr + (int64_t)d * R + r] = f2bf(v);
  }
}

// =====================================================================================================
// Persistent forward kernel.  Roles: warp 0 TMA producer (+ L2 prefetch), warp 1 MMA issuer, warps 2..17 epilogue
// (two groups of 8 warps alternating tiles; (TMEM lane quarter) x (channel half)), warp 18 TMA-store issuer; two CTAs
// per SM fill each other's bubbles.
// Measured on B200 (tools/mma_cost.cu, tools/trace_layer.py): one tcgen05.mma (M = 128, K = 16, N <= 64) occupies the
// tensor pipe for ~48 cycles and its single-thread issue costs about as much again, and an elected epilogue thread
// that issues a TMA store stalls its whole group for ~450 cycles -- hence: as few MMA instructions as possible (6 per
// tile; biases and the residual add stay on the CUDA cores), one-add descriptors, and a warp that does nothing but
// issue stores and hand buffers back.
// =====================================================================================================
struct LayerFwdPArgs {
  const float* params;
  int64_t sig_b, gate_b, res_b;
  int T, dil, dil_next, l, last, n_tiles, tiles_per_slot;
  int z_col;  // first column of this layer's block in the z stash
  int* tile_ctr;  // [2] dynamic tile scheduler: next-tile counter, finished-CTA counter (zero on entry, reset on exit)
  const float* gc_tbl;  // global conditioning: this layer's [C1][2D] projection table (tmodel.py:150-154), else nullptr
  const int32_t* ids;   // [B][T] voice ids
  int C1;
  // local conditioning (tmodel.py:156-160): this layer's plane [B*T][2D] bf16 = lc_upsampled . [LC_SIGNAL | LC_GATE],
  // added to the pre-activations row by row in the gate epilogue
  const bf16* cond;
  uint64_t pol_x0, pol_z, pol_xout;  // L2 eviction hints (umma.cuh): x[t-dil] tile = last forward use of those rows; z; x'
  long long* trace;
};
// 4 bf16 (two packed words) added onto a float4 of pre-activation biases, scaled (GATE half: 0.5, see the file header)
__device__ __forceinline__ void add_bf16x4(float4& b, uint32_t w0, uint32_t w1, float sc) {
  b.x = fmaf(sc, __uint_as_float(w0 << 16), b.x); b.y = fmaf(sc, __uint_as_float(w0 & 0xffff0000u), b.y);
  b.z = fmaf(sc, __uint_as_float(w1 << 16), b.z); b.w = fmaf(sc, __uint_as_float(w1 & 0xffff0000u), b.w);
}

template <int R, int D, bool GC, bool LC>
__global__ void __launch_bounds__(608, 2)
k_layer_fwd_p_umma(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_xout,
                   const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_wc,
                   const __grid_constant__ CUtensorMap map_wr, LayerFwdPArgs a) {
  static_assert(R == 32 && D == 32, "64-byte activation rows");
  constexpr int XB = 64;
  constexpr int PANEL = 128 * XB;       // 8 KB
  constexpr int STAGE = 2 * PANEL;      // x[t-dil] (later: the output tile) | x[t]
  constexpr int NST = 5;
  constexpr uint32_t HI = desc_hi(XB);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* zt = smem + NST * STAGE;       // [2] z tiles, one per epilogue group
  unsigned char* wc0 = zt + 2 * PANEL;          // [2D rows][R] 4 KB (GATE half pre-scaled by 0.5)
  unsigned char* wc1 = wc0 + 2 * D * XB;
  unsigned char* wr = wc1 + 2 * D * XB;         // [R rows][D] 2 KB
  __shared__ __align__(8) uint64_t w_full, in_full[NST], stage_free[NST], v_full[2], r_full[2], v_free[2], r_free[2],
      z_ready[2], zo_ready[2], xo_ready[NST], zt_free[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float bias_s[96];   // SIGNAL_BIAS | 0.5 * GATE_BIAS | RESIDUAL_BIAS
  // Dynamic tile scheduler: the first tile of a CTA is blockIdx.x, further ones come from a global counter (SMs do not
  // run at the same speed and the second CTA of an SM starts ~1.4 us late: a static split leaves a 4-8 us tail).
  // tile_s[stage] = tile held by that ring stage, -1 = no more tiles.  The MMA issuer reads it after in_full[stage];
  // the epilogue groups after v_full (for a sentinel the issuer commits v_full without any MMA); the store warp gets
  // the tile with the z / x' hand-over and the final count from n_done_s.
  __shared__ int tile_s[NST], zo_tile[2], xo_tile[NST];
  __shared__ volatile int n_done_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_launch_dependents();  // the next layer's CTAs may take this SM's resources as soon as they are released
  if (tid == 0) n_done_s = 0x7fffffff;
  LayerTracer tr;
  tr.init(a.trace, warp, blockIdx.x == 0 && lane == 0);
  tr.ev(30, 0);
  if (kLayerTrace && a.trace != nullptr && tid == 0) {  // per-CTA wall-clock start / SM id (tools/trace_layer.py)
    unsigned long long gt;
    unsigned smid;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    a.trace[32 * WN_TRACE_PER_WARP + 4 * blockIdx.x] = (long long)gt;
    a.trace[32 * WN_TRACE_PER_WARP + 4 * blockIdx.x + 2] = (long long)smid;
  }
  if (tid == 0) {
    mbar_init(&w_full, 1);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&in_full[i], 1);
      mbar_init(&stage_free[i], 1);
      mbar_init(&xo_ready[i], 1);  // per ring stage: a group that runs ahead of the store warp cannot lap a phase
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&r_full[i], 1);
      mbar_init(&v_free[i], 256);
      mbar_init(&r_free[i], 256);
      mbar_init(&z_ready[i], 1);
      mbar_init(&zo_ready[i], 1);
      mbar_init(&zt_free[i], 1);
    }
    fence_mbar_init();
  }
  tr.ev(34, 0);
  if (tid < 96)
    bias_s[tid] = tid < 32 ? (a.sig_b >= 0 ? a.params[a.sig_b + tid] : 0.f)
                : tid < 64 ? (a.gate_b >= 0 ? 0.5f * a.params[a.gate_b + tid - 32] : 0.f)
                           : (a.res_b >= 0 ? a.params[a.res_b + tid - 64] : 0.f);
  tr.ev(35, 0);
  if (warp == 1) tmem_alloc(&tmem_base_s, 256);
  tr.ev(36, 0);
  tc_fence_before_sync();
  __syncthreads();
  tr.ev(37, 0);
  tc_fence_after_sync();
  const uint32_t tm = tmem_base_s;  // buffer ab: acc_v at ab*128 (64 cols), acc_r at ab*128 + 64 (32 cols)
  // Programmatic dependent launch: everything above touched only parameters and on-chip state; the previous layer's
  // activations (and the buffers this kernel overwrites) are safe to use once the prerequisite grid has completed.
  pdl_wait();
  tr.ev(31, 0);

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&w_full, (uint32_t)(2 * 2 * D * XB + (a.last ? 0 : R * XB)));
      tma_load_2d(wc0, &map_wc, &w_full, 0, (a.l * 2 + 0) * 2 * D);
      tma_load_2d(wc1, &map_wc, &w_full, 0, (a.l * 2 + 1) * 2 * D);
      if (!a.last) tma_load_2d(wr, &map_wr, &w_full, 0, a.l * R);
      int tile = (int)blockIdx.x, i = 0;
      for (;; ++i) {
        const int s = i % NST;
        tr.ev(1, i);
        mbar_wait(&stage_free[s], ((uint32_t)(i / NST) & 1u) ^ 1u);
        tr.ev(2, i);
        if (tile >= a.n_tiles) break;
        const int b = tile / a.tiles_per_slot, t0 = (tile % a.tiles_per_slot) * 128;
        tile_s[s] = tile;
        mbar_expect_tx(&in_full[s], (uint32_t)STAGE);
        tma_load_3d_hint(smem + s * STAGE, &map_x, &in_full[s], 0, t0, b, a.pol_x0);
        tma_load_3d(smem + s * STAGE + PANEL, &map_x, &in_full[s], 0, t0 + a.dil, b);
        tile = (int)gridDim.x + atomicAdd(a.tile_ctr, 1);  // the next one, fetched while these loads fly
      }
      n_done_s = i;
      for (int k = 0; k < 2; ++k) {  // one sentinel per epilogue group (they alternate tiles)
        const int j = i + k, s = j % NST;
        if (k > 0) mbar_wait(&stage_free[s], ((uint32_t)(j / NST) & 1u) ^ 1u);
        tile_s[s] = -1;
        mbar_arrive(&in_full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(&w_full, 0);
      const uint32_t idv = make_idesc_bf16(128, 2 * D), idr = make_idesc_bf16(128, R);
      const uint32_t ring_lo = desc_lo_k(smem_u32(smem)), zt_lo = desc_lo_k(smem_u32(zt));
      const uint32_t wc0_lo = desc_lo_k(smem_u32(wc0)), wc1_lo = desc_lo_k(smem_u32(wc1)), wr_lo = desc_lo_k(smem_u32(wr));
      auto mma1 = [&](int i) {  // conv taps: acc_v = x[t-dil] . W[0] + x[t] . W[1]   (K = 16 per instruction: 32 bytes)
        const int s = i % NST, ab = i & 1;
        const uint32_t x0 = ring_lo + (uint32_t)s * (STAGE >> 4), x1 = x0 + (PANEL >> 4), av = tm + ab * 128;
        mma_bf16_ss2(av, x0, HI, wc0_lo, HI, idv, false);
        mma_bf16_ss2(av, x0 + 2, HI, wc0_lo + 2, HI, idv, true);
        mma_bf16_ss2(av, x1, HI, wc1_lo, HI, idv, true);
        mma_bf16_ss2(av, x1 + 2, HI, wc1_lo + 2, HI, idv, true);
        mma_commit(&v_full[ab]);
      };
      auto mma2 = [&](int i) {  // residual 1x1: acc_r = z . RESIDUAL
        const int ab = i & 1;
        const uint32_t z = zt_lo + (uint32_t)ab * (PANEL >> 4), ar = tm + ab * 128 + 64;
        mma_bf16_ss2(ar, z, HI, wr_lo, HI, idr, false);
        mma_bf16_ss2(ar, z + 2, HI, wr_lo + 2, HI, idr, true);
        mma_commit(&r_full[ab]);
      };
      // two queues served in whatever order their inputs become ready
      int n1 = 0, n2 = 0, n_end = 1 << 30;  // n_end: index of the first sentinel
      uint32_t spins = 0;
      while (n1 < n_end + 2 || n2 < n_end) {
        bool did = false;
        if (a.last) n2 = min(n1, n_end);
        if (n2 < min(n1, n_end) && mbar_test_wait(&z_ready[n2 & 1], (uint32_t)(n2 >> 1) & 1u) &&
            mbar_test_wait(&r_free[n2 & 1], ((uint32_t)(n2 >> 1) & 1u) ^ 1u)) {
          tc_fence_after_sync();
          tr.ev(4, n2);
          mma2(n2++);
          did = true;
        }
        if (n1 < n_end + 2 && mbar_test_wait(&in_full[n1 % NST], (uint32_t)(n1 / NST) & 1u) &&
            mbar_test_wait(&v_free[n1 & 1], ((uint32_t)(n1 >> 1) & 1u) ^ 1u)) {
          tc_fence_after_sync();
          if (tile_s[n1 % NST] < 0) {  // sentinel: wake the owning epilogue group without any MMA
            if (n_end == (1 << 30)) n_end = n1;
            mma_commit(&v_full[n1 & 1]);
            ++n1;
          } else {
            tr.ev(3, n1);
            mma1(n1++);
          }
          did = true;
        }
        if (did) spins = 0; else if (++spins > (1u << 26)) __trap();
      }
    }
  } else if (warp < 18) {
    // ===== epilogue: two groups of 8 warps that alternate tiles (group g owns tiles g, g+2, ... and with them the
    // accumulator parity g and z tile g); one group's TMEM round trip / proxy fence / barrier / wait for the residual
    // MMA overlaps the other group's arithmetic.  thread <-> (row r, channels [16*half, +16)), two passes of 8 =====
    const int e = warp - 2;
    const int g = e >> 3, half = (e >> 2) & 1, q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const bool elected = ((e & 7) == 0 && lane == 0);
    const uint32_t sw = ((uint32_t)r >> 1) & 3u;  // SW64: 16-byte chunk index ^= (row / 2) % 4
    const uint32_t o[2] = {(uint32_t)r * 64u + ((((uint32_t)(2 * half)) ^ sw) << 4),
                           (uint32_t)r * 64u + ((((uint32_t)(2 * half + 1)) ^ sw) << 4)};
    const float4* bs4 = reinterpret_cast<const float4*>(bias_s + 16 * half);
    const float4* bg4 = reinterpret_cast<const float4*>(bias_s + 32 + 16 * half);
    const float4* br4 = reinterpret_cast<const float4*>(bias_s + 64 + 16 * half);
    unsigned char* ztile = zt + g * PANEL;
    const uint32_t tb = tm + g * 128 + lane_sel;
    auto gbar = [&]() {
      if (g == 0) asm volatile("bar.sync 1, 256;" ::: "memory"); else asm volatile("bar.sync 2, 256;" ::: "memory");
    };
    for (int i = g;; i += 2) {
      const int s = i % NST, ab = g;
      // ---- gate: z = tanh(v_s + b_s) * sigmoid(v_g + b_g) ----
      tr.ev(5, i);
      uint4 lcs[2], lcg[2];  // this row's local-conditioning projections (signal / gate channels of the two passes)
      if constexpr (LC) {
        // the tile id is published before the stage's loads are issued: visible once in_full[s] has completed (the stage
        // cannot be recycled before this tile's epilogue has run).  The plane row is fetched while the conv MMA runs.
        mbar_wait(&in_full[s], (uint32_t)(i / NST) & 1u);
        const int tl = tile_s[s];
        lcs[0] = lcs[1] = lcg[0] = lcg[1] = make_uint4(0u, 0u, 0u, 0u);
        if (tl >= 0) {
          const int b = tl / a.tiles_per_slot, t = (tl % a.tiles_per_slot) * 128 + r;
          if (t < a.T) {
            const uint4* row = reinterpret_cast<const uint4*>(a.cond + ((size_t)b * a.T + t) * (2 * D) + 16 * half);
            lcs[0] = __ldg(row); lcs[1] = __ldg(row + 1);
            lcg[0] = __ldg(row + D / 8); lcg[1] = __ldg(row + D / 8 + 1);
          }
        }
      }
      mbar_wait(&v_full[ab], (uint32_t)(i >> 1) & 1u);
      tr.ev(6, i);
      const int tile = tile_s[s];  // written before the loads the conv MMA (or the sentinel commit) waited for
      if (tile < 0) break;
      tc_fence_after_sync();
      const float* gct = nullptr;  // this row's global-conditioning projections (signal [D] | gate [D])
      if constexpr (GC) {
        const int b = tile / a.tiles_per_slot, t = (tile % a.tiles_per_slot) * 128 + r;
        const int id = t < a.T ? min(max(__ldg(a.ids + (size_t)b * a.T + t), 0), a.C1 - 1) : 0;
        gct = a.gc_tbl + (size_t)id * 2 * D + 16 * half;
      }
      if (i >= 2) mbar_wait(&zt_free[ab], (uint32_t)((i - 2) >> 1) & 1u);  // z(i-2)'s store has finished reading zt[g]
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        uint32_t vs[8], vg[8], pk[4];
        tmem_ld_32x32b_x8(tb + 16 * half + 8 * p, vs);
        tmem_ld_32x32b_x8(tb + 32 + 16 * half + 8 * p, vg);
        tmem_ld_wait();
        if (p == 1) {
          tc_fence_before_sync();
          mbar_arrive(&v_free[ab]);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float4 b_s = bs4[2 * p + q], b_g = bg4[2 * p + q];
          if constexpr (GC) {  // the accumulator's GATE half holds 0.5 * pre-activation: the table's gate half is halved too
            const float4 c_s = __ldg(reinterpret_cast<const float4*>(gct) + 2 * p + q);
            const float4 c_g = __ldg(reinterpret_cast<const float4*>(gct + D) + 2 * p + q);
            b_s.x += c_s.x; b_s.y += c_s.y; b_s.z += c_s.z; b_s.w += c_s.w;
            b_g.x = fmaf(0.5f, c_g.x, b_g.x); b_g.y = fmaf(0.5f, c_g.y, b_g.y);
            b_g.z = fmaf(0.5f, c_g.z, b_g.z); b_g.w = fmaf(0.5f, c_g.w, b_g.w);
          }
          if constexpr (LC) {
            const uint32_t ws0 = q == 0 ? lcs[p].x : lcs[p].z, ws1 = q == 0 ? lcs[p].y : lcs[p].w;
            const uint32_t wg0 = q == 0 ? lcg[p].x : lcg[p].z, wg1 = q == 0 ? lcg[p].y : lcg[p].w;
            add_bf16x4(b_s, ws0, ws1, 1.f);
            add_bf16x4(b_g, wg0, wg1, 0.5f);
          }
          const float z0 = tanh_fast(__uint_as_float(vs[4 * q]) + b_s.x) * fmaf(0.5f, tanh_fast(__uint_as_float(vg[4 * q]) + b_g.x), 0.5f);
          const float z1 = tanh_fast(__uint_as_float(vs[4 * q + 1]) + b_s.y) * fmaf(0.5f, tanh_fast(__uint_as_float(vg[4 * q + 1]) + b_g.y), 0.5f);
          const float z2 = tanh_fast(__uint_as_float(vs[4 * q + 2]) + b_s.z) * fmaf(0.5f, tanh_fast(__uint_as_float(vg[4 * q + 2]) + b_g.z), 0.5f);
          const float z3 = tanh_fast(__uint_as_float(vs[4 * q + 3]) + b_s.w) * fmaf(0.5f, tanh_fast(__uint_as_float(vg[4 * q + 3]) + b_g.w), 0.5f);
          pk[2 * q] = pack2(z0, z1);
          pk[2 * q + 1] = pack2(z2, z3);
        }
        *reinterpret_cast<uint4*>(ztile + o[p]) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async_smem();
      tr.ev(7, i);
      gbar();
      tr.ev(8, i);
      if (elected) {
        zo_tile[ab] = tile;  // (the ring stage may have been recycled by the time the store warp gets here: last layer)
        xo_tile[s] = tile;
        if (!a.last) mbar_arrive(&z_ready[ab]);
        else mbar_arrive(&stage_free[s]);  // last layer: both x tiles were only read by the (completed) conv MMA
        mbar_arrive(&zo_ready[ab]);
      }
      if (a.last) continue;
      // ---- x' = bf16(x[t] + z . RESIDUAL + bias), written over the dead x[t-dil] tile ----
      unsigned char* otile = smem + s * STAGE;
      const unsigned char* x1 = otile + PANEL;
      tr.ev(9, i);
      mbar_wait(&in_full[s], (uint32_t)(i / NST) & 1u);  // x[t] tile (TMA) visible to this thread
      mbar_wait(&r_full[ab], (uint32_t)(i >> 1) & 1u);
      tr.ev(10, i);
      tc_fence_after_sync();
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        uint32_t vr[8], pk[4];
        tmem_ld_32x32b_x8(tb + 64 + 16 * half + 8 * p, vr);
        const uint4 xa = *reinterpret_cast<const uint4*>(x1 + o[p]);
        const uint32_t xin[4] = {xa.x, xa.y, xa.z, xa.w};
        tmem_ld_wait();
        if (p == 1) {
          tc_fence_before_sync();
          mbar_arrive(&r_free[ab]);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float4 b_r = br4[2 * p + q];
          pk[2 * q] = pack2(__uint_as_float(vr[4 * q]) + __uint_as_float(xin[2 * q] << 16) + b_r.x,
                            __uint_as_float(vr[4 * q + 1]) + __uint_as_float(xin[2 * q] & 0xffff0000u) + b_r.y);
          pk[2 * q + 1] = pack2(__uint_as_float(vr[4 * q + 2]) + __uint_as_float(xin[2 * q + 1] << 16) + b_r.z,
                                __uint_as_float(vr[4 * q + 3]) + __uint_as_float(xin[2 * q + 1] & 0xffff0000u) + b_r.w);
        }
        *reinterpret_cast<uint4*>(otile + o[p]) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async_smem();
      tr.ev(11, i);
      gbar();
      tr.ev(12, i);
      if (elected) mbar_arrive(&xo_ready[s]);
    }
  } else if (warp == 18) {
    // ===== TMA-store issuer: serves the z and x' queues in whatever order they fill; hands zt buffers and ring stages back =====
    if (lane == 0) {
      int prev_kind = -1, prev_idx = 0;  // newest committed bulk group: 0 = z tile, 1 = x' tile
      auto release_prev = [&]() {
        if (prev_kind < 0) return;
        tma_store_wait_read<1>();  // every group but the newest has finished reading shared memory
        if (prev_kind == 0) mbar_arrive(&zt_free[prev_idx & 1]);
        else mbar_arrive(&stage_free[prev_idx % NST]);
      };
      int nz = 0, nx = 0;
      uint32_t spins = 0;
      while (nz < n_done_s || (!a.last && nx < n_done_s)) {  // n_done_s: INT_MAX until the producer has run out of tiles
        bool did = false;
        if (nz < n_done_s && mbar_test_wait(&zo_ready[nz & 1], (uint32_t)(nz >> 1) & 1u)) {
          const int tile = zo_tile[nz & 1];
          const int b = tile / a.tiles_per_slot, t0 = (tile % a.tiles_per_slot) * 128;
          tma_store_3d_hint(&map_z, zt + (nz & 1) * PANEL, a.z_col, t0, b, a.pol_z);
          tma_store_commit();
          release_prev();
          prev_kind = 0; prev_idx = nz++;
          did = true;
        }
        // (last layer: no x' tiles at all -- xo_ready never completes a phase, and a parity test on a barrier one is not
        // following phase by phase gives false positives)
        if (!a.last && nx < nz && mbar_test_wait(&xo_ready[nx % NST], (uint32_t)(nx / NST) & 1u)) {
          const int tile = xo_tile[nx % NST];
          const int b = tile / a.tiles_per_slot, t0 = (tile % a.tiles_per_slot) * 128;
          tma_store_3d_hint(&map_xout, smem + (nx % NST) * STAGE, 0, a.dil_next + t0, b, a.pol_xout);
          tma_store_commit();
          tr.ev(13, nx);
          release_prev();
          prev_kind = 1; prev_idx = nx++;
          did = true;
        }
        if (did) spins = 0; else if (++spins > (1u << 26)) __trap();
      }
      tma_store_wait_all<0>();
    }
  }
  tr.ev(32, 0);
  tc_fence_before_sync();
  __syncthreads();
  tr.ev(33, 0);
  if (tid == 0 && atomicAdd(a.tile_ctr + 1, 1) == (int)gridDim.x - 1) {  // last CTA out re-arms the scheduler
    a.tile_ctr[0] = 0;
    a.tile_ctr[1] = 0;
  }
  if (kLayerTrace && a.trace != nullptr && tid == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    a.trace[32 * WN_TRACE_PER_WARP + 4 * blockIdx.x + 1] = (long long)gt;
  }
  if (warp == 1) tmem_dealloc(tm, 256);
}

struct LayerBwdFusedArgs {
  const float* params;
  float* grads;
  int64_t sig, gate, res, sig_b, gate_b, res_b;
  int dil, l, has_next, n_tiles, tiles_per_slot, z_plane0;
  int n_stages;  // ring depth (what fits beside the carry slots: 6 for dil <= 128)
  int n_later;   // ceil(dil / 128): how many later tiles' P0 a tile's outputs can reach into
  int n_carry;   // n_later + 2 carry slots
  const float* gc_tbl;  // global conditioning: this layer's [C1][2D] projection table, else nullptr
  float* dgc_tbl;       // its gradient (fp32 atomics)
  const int32_t* ids;   // [B][T] voice ids
  int C1, T;
  // local conditioning: this layer's plane [B*T][2D] bf16, the conditioning term of the recomputed pre-activations, and
  // the plane of the same shape that receives dv = [dv_s | dv_g], the gradient wrt it (consumed by the LC weight / data
  // gradients of phase L + 1).  Two buffers: a warm-up tile re-reads conditioning rows that belong to another CTA's run.
  const bf16* cond;
  bf16* dcond;
  uint64_t pol_x0, pol_in, pol_out;  // L2 eviction hints: x[t-dil] tile (last use of those rows), dx_{l+1} in, dx_l out
  const bf16* dz;  // this layer's plane of the skip-path gradient, [B * T][D] (read straight from global memory by E1)
  int seq;  // launch sequence number while tracing (tools/trace_layer.py gaps)
  long long* trace;
};

// =====================================================================================================
// k_layer_bwd_fused_umma (see the file header).  One CTA per SM, 28 warps:
//   warp 0        TMA producer                                  warps 1, 27   MMA issuers (queue A / queue B)
//   warps 2..17   gate epilogue E1, two alternating groups      warps 18..25  row warps: E2 (outputs)
//   warp 26       TMA-store issuer (returns ring stages)
// The first version of this kernel (round 1: 40 KB stages X0 | X1 | YN->DX | PN->ONES | DZ, four of them, and an E0
// phase that merged YN + PN into one bf16 tile) was bound by ring depth x stage lifetime: in its in-kernel timeline a
// stage lived ~12 000 cycles, so a tile left the CTA every ~3100 cycles, with HBM at 3.0 of 6.5 TB/s, the tensor pipe
// 28 % and the issue slots 50 % busy; loads landed -> E0 (1200) -> wait (1100) -> queue A -> E1 -> queue B -> E2 -> store
// was the chain.  Hence (1.705 -> 1.50 ms for the 30 layers of configs[1], tile period ~2600 cycles):
//   * the ring stage is X0 | X1 | YN | PN only (32 KB, 5 deep).  dz never goes through shared memory: every E1 thread
//     fetches its own 2 x 16 bytes of the dz plane with plain global loads issued before it waits for the tile's MMAs;
//   * there is no E0 phase: dx_{l+1} = YN + PN is never materialised.  By linearity acc_d = YN . Wr^T + PN . Wr^T (two
//     more N = 32 MMAs) and the MN-major A operand of the weight-gradient MMA is the stage itself, [X0 | X1 | YN | PN]^T:
//     rows 64..95 and 96..127 of acc_w are the two halves of RESIDUAL's (transposed) gradient, added in the flush;
//   * with the constant-one panel gone, the SIGNAL / GATE bias gradients (column sums of dv) come from a third MMA queue
//     (C, issued by queue A's thread in its idle time): [DVs | DVg | ..]^T . 1 with a 1 KB all-ones B operand (N = 16),
//     accumulated in 16 more TMEM columns.  (Summing them in E1 with warp shuffles cost 0.16 ms per step: E1's
//     arithmetic is on the critical chain, the tensor pipe is not.);
//   * the outputs overwrite YN (Y_l) and PN (P0_l) once every MMA reading the stage has completed.
// Work buffer (24 KB, 2 deep): DVs | DVg | Z -- three [128 x 32] SW64 panels written by E1: K-major A operands of
// dv . W^T and, re-described MN-major, the B operand (N = 96) of the weight-gradient MMA and the A operand of queue C.
// Every tcgen05.mma (M = 128, K = 16) re-reads its 4 KB A slice from shared memory, ~48 cycles at N <= 64 whatever N is
// (tools/mma_cost.cu).  28 instructions per tile:
//   A: acc_v (N=64) = X0.W0 + X1.W1 [4] ; acc_d (N=32) = YN . RESIDUAL^T + PN . RESIDUAL^T [4]
//   B: acc_p (N=64) = [DVs|DVg] . [W0^T|W1^T] [4] ; acc_w (N=96) += [X0|X1|YN|PN]^T . [DVs|DVg|Z] [8]
//   C: acc_b (N=16) += [DVs|DVg|Z|..]^T . 1 [8]
// TMEM: per tile parity ab: acc_v [ab*160, +64), acc_d [+64, +32), acc_p [+96, +64); persistent acc_w [320, +96),
// acc_b [416, +16).  RESIDUAL_BIAS gradient = column sums of dx_{l+1}, kept in the row warps' registers.
// Tried on this version and dropped: dz requested a whole tile ahead from DRAM into registers (the pending loads made
// E1's arithmetic ~1000 cycles slower; now the producer L2-prefetches the dz tile and E1 loads it at the top of its
// tile); outputs written with plain 16-byte global stores from E2 instead of shared memory + TMA (the stage is free
// ~3000 cycles earlier, but a warp's 32 rows x 16 bytes are 32 separate sectors: 1.50 -> 1.70 ms).
// =====================================================================================================
// 8 values x 32 lanes -> every lane L returns the warp total of value (L >> 2) & 7 (9 shuffles instead of 8 warp sums)
__device__ __forceinline__ float warp_transpose_sum8(float (&w)[8], int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool hi = (lane & 16) != 0;
    const float send = hi ? w[i] : w[i + 4], keep = hi ? w[i + 4] : w[i];
    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool hi = (lane & 8) != 0;
    const float send = hi ? w[i] : w[i + 2], keep = hi ? w[i + 2] : w[i];
    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    const bool hi = (lane & 4) != 0;
    const float send = hi ? w[0] : w[1], keep = hi ? w[1] : w[0];
    w[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  w[0] += __shfl_xor_sync(0xffffffffu, w[0], 2);
  w[0] += __shfl_xor_sync(0xffffffffu, w[0], 1);
  return w[0];
}

template <int R, int D, bool GC, bool LC>
__global__ void __launch_bounds__(896, 1)
k_layer_bwd_fused_umma(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dz,
                       const __grid_constant__ CUtensorMap map_dxn, const __grid_constant__ CUtensorMap map_dxo,
                       const __grid_constant__ CUtensorMap map_wc, const __grid_constant__ CUtensorMap map_wrn,
                       LayerBwdFusedArgs a) {
  static_assert(R == 32 && D == 32, "tile bookkeeping below assumes 64-byte activation rows");
  constexpr int XB = 64;
  constexpr int PANEL = 128 * XB;                 // 8 KB: one [128 x 32] bf16 tile
  constexpr int P_X0 = 0, P_X1 = 1, P_DX = 2;
  constexpr int STAGE = 3 * PANEL;                // 24 KB
  constexpr int MAX_NST = 6;
  const int NST = a.n_stages;                     // 6 for dil <= 128 (host: what fits beside the carry slots)
  constexpr int W_DVS = 0, W_DVG = 1, W_Z = 2, W_ONES = 3;
  constexpr int WBUF = 4 * PANEL;                 // 32 KB: DVs | DVg | Z | constant 1.0
  constexpr int NE1 = 512, NE1G = 256, NE2 = 256;
  constexpr uint32_t ACC_V = 0, ACC_D = 64, ACC_P = 96, ACC_STRIDE = 160, ACC_W = 320, ACC_B = 432;
  constexpr uint32_t HI = desc_hi(XB);
  constexpr int STG_WC = 0, STG_WR = 64 * 65;  // end-of-kernel staging (floats): 21 KB, inside stage 0
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* wb = smem + NST * STAGE;         // [2] work buffers
  unsigned char* wc0 = wb + 2 * WBUF;             // [2D rows][R]   4 KB (GATE half pre-scaled by 0.5)
  unsigned char* wc1 = wc0 + 2 * D * XB;          //                4 KB, directly behind wc0
  unsigned char* wrn = wc1 + 2 * D * XB;          // [D rows][R]    2 KB
  unsigned char* bt = wrn + D * XB;               // [64 rows][16] bf16, K-major SW32: SIGNAL_BIAS | 0.5 GATE_BIAS as hi + lo
  unsigned char* idn = bt + 2048;                 // [64 rows][32] bf16, K-major: rows 0..31 zero, rows 32..63 the identity
  unsigned char* carry = idn + 4096;              // [n_carry] P0 tiles (bf16, SW64 rows) of this and the later tiles
  float* stg = reinterpret_cast<float*>(smem);    // aliases stage 0
  __shared__ __align__(8) uint64_t w_full, in_full[MAX_NST], stage_free[MAX_NST], out_ready[MAX_NST], v_full[2],
      acc1_free[2], dv_ready[2], p_full[2], acc2_free[2], g_full;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // Work list of this CTA: a contiguous run [j0, j1) of the tiles in (slot, time DEscending) order, preceded by up to
  // n_later warm-up tiles (the tiles just later in time than the first one: only their P0 is wanted).
  const int TPS = a.tiles_per_slot;
  const int j0 = (int)(((long long)blockIdx.x * a.n_tiles) / gridDim.x), j1 = (int)(((long long)(blockIdx.x + 1) * a.n_tiles) / gridDim.x);
  const int k_first = TPS - 1 - j0 % TPS;
  const int n_warm = j1 > j0 ? min(a.n_later, TPS - 1 - k_first) : 0;
  const int n_my = n_warm + (j1 - j0);
  // every role walks the same item list: (slot, time tile k, warm-up?), time running backwards inside a slot
  struct Item {
    int slot, k, warm_left, tps;
    __device__ __forceinline__ bool warm() const { return warm_left > 0; }
    __device__ __forceinline__ void next() {
      if (warm_left > 0) { --warm_left; --k; }
      else if (k == 0) { k = tps - 1; ++slot; }
      else --k;
    }
  };
  const Item item0{j0 / TPS, k_first + n_warm, n_warm, TPS};
  pdl_launch_dependents();
  if (kLayerTrace && a.trace != nullptr && tid == 0) {  // kernel entry (before barrier init / TMEM allocation)
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    a.trace[32 * WN_TRACE_PER_WARP + 4 * blockIdx.x + 3] = (long long)gt;
    atomicMin(reinterpret_cast<unsigned long long*>(a.trace) + 32 * WN_TRACE_PER_WARP + 4096 + 2 * (a.seq & 63), gt);
  }

  if (tid == 0) {
    mbar_init(&w_full, 1);
    mbar_init(&g_full, 2);
    for (int i = 0; i < MAX_NST; ++i) {
      mbar_init(&in_full[i], 1);
      mbar_init(&stage_free[i], 1);
      mbar_init(&out_ready[i], 8);  // one arrival per row warp
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&acc1_free[i], NE1G);
      mbar_init(&dv_ready[i], 1);
      mbar_init(&p_full[i], 2);  // queue B + queue C
      mbar_init(&acc2_free[i], NE2);
    }
    fence_mbar_init();
  }
  // the pre-activation biases enter acc_v as one more K = 16 step, (constant ones) x (bias as hi + lo bf16, 2^-17
  // relative): 32 adds and 8 shared-memory loads less per gate-epilogue thread and tile, where issue slots are scarce
  build_bias_tile(bt, 64, tid, 896, [&](int n) {
    return n < 32 ? (a.sig_b >= 0 ? a.params[a.sig_b + n] : 0.f) : (a.gate_b >= 0 ? 0.5f * a.params[a.gate_b + n - 32] : 0.f);
  });
  fill_ones(wb + W_ONES * PANEL, PANEL, tid, 896);
  fill_ones(wb + WBUF + W_ONES * PANEL, PANEL, tid, 896);
  if (tid < 256) {  // [0 | I]: contracted against the DX tile it adds dx_{l+1} onto the P1 half of acc_p (exact: 1.0 x bf16)
    const int n = tid >> 2, ch = tid & 3;
    const int e = (n - 32) & 7, wi = e >> 1;
    const bool hit = n >= 32 && ch == ((n - 32) >> 3);
    const uint32_t one = 0x3f80u << (16 * (e & 1));
    *reinterpret_cast<uint4*>(idn + swizzled_offset((uint32_t)n, (uint32_t)(ch * 16), 64)) =
        make_uint4(hit && wi == 0 ? one : 0u, hit && wi == 1 ? one : 0u, hit && wi == 2 ? one : 0u, hit && wi == 3 ? one : 0u);
  }
  fence_proxy_async_smem();

  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = tmem_base_s;
  pdl_wait();  // the layer above has finished writing (Y, P0) and reading the buffers this layer overwrites
  LayerTracer tr;
  tr.init(a.trace, warp, blockIdx.x == 0 && lane == 0);
  if (kLayerTrace && a.trace != nullptr && tid == 0) {  // per-CTA wall-clock start / SM id (tools/trace_layer.py)
    unsigned long long gt;
    unsigned smid;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    a.trace[32 * WN_TRACE_PER_WARP + 4 * blockIdx.x] = (long long)gt;
    a.trace[32 * WN_TRACE_PER_WARP + 4 * blockIdx.x + 2] = (long long)smid;
  }

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      mbar_expect_tx(&w_full, (uint32_t)(2 * 2 * D * XB + D * XB));
      tma_load_2d(wc0, &map_wc, &w_full, 0, (a.l * 2 + 0) * 2 * D);
      tma_load_2d(wc1, &map_wc, &w_full, 0, (a.l * 2 + 1) * 2 * D);
      tma_load_2d(wrn, &map_wrn, &w_full, 0, a.l * D);
      Item it = item0;
      for (int i = 0; i < n_my; ++i, it.next()) {
        const int b = it.slot, t0 = it.k * 128;
        const int s = i % NST;
        unsigned char* st = smem + s * STAGE;
        tr.ev(1, i);
        mbar_wait(&stage_free[s], ((uint32_t)(i / NST) & 1u) ^ 1u);
        tr.ev(2, i);
        mbar_expect_tx(&in_full[s], (uint32_t)((a.has_next ? 3 : 2) * PANEL));
        tma_prefetch_l2_3d(&map_dz, 0, t0, a.z_plane0 + b);  // E1 reads this tile's dz rows from the L2 a few thousand cycles from now
        tma_load_3d_hint(st + P_X0 * PANEL, &map_x, &in_full[s], 0, t0, b, a.pol_x0);
        tma_load_3d(st + P_X1 * PANEL, &map_x, &in_full[s], 0, t0 + a.dil, b);
        if (a.has_next) tma_load_3d_hint(st + P_DX * PANEL, &map_dxn, &in_full[s], 0, t0, b, a.pol_in);
      }
    }
  } else if (warp == 1 || warp == 27) {
    // ===== MMA issuers: warp 1 serves queue A, warp 27 queue B.  One thread spends ~100 cycles per tcgen05.mma it issues
    // (descriptor arithmetic, the instruction itself, polling) against ~50 cycles of tensor-pipe time.  tcgen05.commit
    // tracks the issuing thread's own MMAs, and the two queues write disjoint accumulators, so they only meet through the
    // mbarriers. =====
    if (lane == 0) {
      mbar_wait(&w_full, 0);
      const uint32_t idv = make_idesc_bf16(128, 2 * D), idd = make_idesc_bf16(128, D);
      const uint32_t idp = make_idesc_bf16(128, 2 * R, false, true);  // A = dv (K-major), B = [W0|W1] (MN-major)
      // B = [DVs | DVg | Z | 1] (N = 112): the ones columns make column 96 of acc_w the column sums of the A operand, whose
      // rows 64..95 (dx_{l+1}) are the RESIDUAL_BIAS gradient -- no instruction, no register, no epilogue time spent on it
      const uint32_t idw = make_idesc_bf16(128, 2 * D + R + 16, true, true);
      // K-major operands: K step of 16 elements = 32 bytes = +2 in the descriptor; MN-major: 16 rows of 64 bytes = +64
      const uint32_t ring_k = desc_lo_k(smem_u32(smem)), ring_mn = desc_lo(smem_u32(smem), PANEL);
      const uint32_t wb_k = desc_lo_k(smem_u32(wb)), wb_mn = desc_lo(smem_u32(wb), PANEL);
      const uint32_t wc0_k = desc_lo_k(smem_u32(wc0)), wc1_k = desc_lo_k(smem_u32(wc1));
      const uint32_t wc_mn = desc_lo(smem_u32(wc0), 2 * D * XB);  // chunk 0 = wc0 (-> P0), chunk 1 = wc1 (-> Y)
      const uint32_t wrn_k = desc_lo_k(smem_u32(wrn)), idn_k = desc_lo_k(smem_u32(idn));
      // queue C: A = the work buffer re-described MN-major with M = 128 (rows 0..63 = dv channels; 64..95 = z and 96..127
      // = the ones panel: never read back), B = 16 rows of the ones panel of work buffer 0, the same for every K step
      const uint32_t idb = make_idesc_bf16(128, 16, true, true), ones_mn = desc_lo(smem_u32(wb + W_ONES * PANEL), PANEL);
      const uint32_t ones_k = desc_lo_k(smem_u32(wb + W_ONES * PANEL)), bt_k = desc_lo_k(smem_u32(bt));
      constexpr uint32_t HI32 = desc_hi(32);
      auto issue_a = [&](int i) {  // recomputed pre-activations; residual part of dz
        const int s = i % NST, ab = i & 1;
        const uint32_t x0 = ring_k + (uint32_t)s * (STAGE >> 4), x1 = x0 + (PANEL >> 4);
        const uint32_t dx = x0 + P_DX * (PANEL >> 4);
        const uint32_t av = tm + ab * ACC_STRIDE + ACC_V, ad = tm + ab * ACC_STRIDE + ACC_D;
        mma_bf16_ss2(av, x0, HI, wc0_k, HI, idv, false);
        mma_bf16_ss2(av, x0 + 2, HI, wc0_k + 2, HI, idv, true);
        mma_bf16_ss2(av, x1, HI, wc1_k, HI, idv, true);
        mma_bf16_ss2(av, x1 + 2, HI, wc1_k + 2, HI, idv, true);
        mma_bf16_ss2(av, ones_k, HI, bt_k, HI32, idv, true);  // + biases
        if (a.has_next) {  // dz(res) = dx_{l+1} . RESIDUAL^T : B = RESIDUAL [D rows][R]
          mma_bf16_ss2(ad, dx, HI, wrn_k, HI, idd, false);
          mma_bf16_ss2(ad, dx + 2, HI, wrn_k + 2, HI, idd, true);
        }
        mma_commit(&v_full[ab]);
      };
      bool acc_started = false;  // the persistent accumulators (acc_w / acc_b) start with this thread's first real tile
      auto issue_b = [&](int i) {  // data gradient + (unless it is a warm-up tile) weight gradients of tile i
        const int s = i % NST, ab = i & 1;
        const uint32_t wk = wb_k + (uint32_t)ab * (WBUF >> 4), wm = wb_mn + (uint32_t)ab * (WBUF >> 4);
        const uint32_t sm = ring_mn + (uint32_t)s * (STAGE >> 4);
        const uint32_t ap = tm + ab * ACC_STRIDE + ACC_P;
        // [P0 | P1] = dv . [W0^T | W1^T]: K = 2D dv channels; B rows 16j.. of the stacked [2D][R] filter copies
        mma_bf16_ss2(ap, wk + W_DVS * (PANEL >> 4), HI, wc_mn, HI, idp, false);
        mma_bf16_ss2(ap, wk + W_DVS * (PANEL >> 4) + 2, HI, wc_mn + 64, HI, idp, true);
        mma_bf16_ss2(ap, wk + W_DVG * (PANEL >> 4), HI, wc_mn + 128, HI, idp, true);
        mma_bf16_ss2(ap, wk + W_DVG * (PANEL >> 4) + 2, HI, wc_mn + 192, HI, idp, true);
        if (a.has_next && i >= n_warm) {  // P1 += dx_{l+1} . [0 | I]: the row warps never touch the DX tile
          const uint32_t dx = ring_k + (uint32_t)s * (STAGE >> 4) + P_DX * (PANEL >> 4);
          mma_bf16_ss2(ap, dx, HI, idn_k, HI, idv, true);
          mma_bf16_ss2(ap, dx + 2, HI, idn_k + 2, HI, idv, true);
        }
        if (i >= n_warm) {
          // A = the stage re-described MN-major with M = 128: rows 0..63 = x[t-dil] | x[t], 64..95 = dx_{l+1}, 96..127 =
          // whatever follows the stage (never read back)
#pragma unroll
          for (int k = 0; k < 8; ++k)  // K = 128 timesteps, 16 per instruction
            mma_bf16_ss2(tm + ACC_W, sm + k * 64, HI, wm + k * 64, HI, idw, acc_started || k != 0);
          acc_started = true;
        }
        mma_commit(&p_full[ab]);
      };
      auto issue_c = [&](int i) {  // column sums of dv over the tile's 128 timesteps: [DVs|DVg|Z|..]^T . 1
        const int ab = i & 1;
        const uint32_t wm = wb_mn + (uint32_t)ab * (WBUF >> 4);
        if (i >= n_warm) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            mma_bf16_ss2(tm + ACC_B, wm + k * 64, HI, ones_mn, HI, idb, acc_started || k != 0);
          acc_started = true;
        }
        mma_commit(&p_full[ab]);
      };
      uint32_t spins = 0;
      if (warp == 27) {
        for (int nb = 0; nb < n_my;) {
          // tile nb's dv tile is written after its v_full, i.e. after queue A issued tile nb
          if (mbar_test_wait(&dv_ready[nb & 1], (uint32_t)(nb >> 1) & 1u) &&
              mbar_test_wait(&acc2_free[nb & 1], ((uint32_t)(nb >> 1) & 1u) ^ 1u)) {
            tc_fence_after_sync();
            tr.ev(4, nb);
            issue_b(nb++);
            tr.ev(17, nb - 1);
            spins = 0;
          } else if (++spins > (1u << 26)) __trap();
        }
        mma_commit(&g_full);
      } else {
        // queue A first (it feeds E1, the longest phase); queue C (bias-gradient column sums) in its shadow
        for (int na = 0, nc = 0; nc < n_my;) {
          bool did = false;
          if (na < n_my && mbar_test_wait(&in_full[na % NST], (uint32_t)(na / NST) & 1u) &&
              mbar_test_wait(&acc1_free[na & 1], ((uint32_t)(na >> 1) & 1u) ^ 1u)) {
            tc_fence_after_sync();
            tr.ev(3, na);
            issue_a(na++);
            tr.ev(16, na - 1);
            did = true;
          }
          if (nc < na && mbar_test_wait(&dv_ready[nc & 1], (uint32_t)(nc >> 1) & 1u)) {
            tc_fence_after_sync();
            issue_c(nc++);
            did = true;
          }
          if (did) spins = 0; else if (++spins > (1u << 26)) __trap();
        }
        mma_commit(&g_full);
      }
    }
  } else if (warp < 18) {
    // ===== E1: gate backward, two groups of 8 warps that alternate tiles (group g owns tiles g, g+2, ... and with them
    // the accumulator / work-buffer parity g), so one group's TMEM round trip, p_full wait, proxy fence and barrier
    // overlap the other group's arithmetic.  thread <-> (row r, channels [16*half, +16)), two passes of 8 =====
    const int e = warp - 2;
    const int g = e >> 3, half = (e >> 2) & 1, q4 = warp & 3;
    const int r = q4 * 32 + lane;
    const int et = e * 32 + lane;  // 0..511
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const bool elected = ((e & 7) == 0 && lane == 0);
    const uint32_t sw64 = ((uint32_t)r >> 1) & 3u;
    const uint32_t o[2] = {(uint32_t)r * 64u + ((((uint32_t)(2 * half)) ^ sw64) << 4),
                           (uint32_t)r * 64u + ((((uint32_t)(2 * half + 1)) ^ sw64) << 4)};
    Item it = item0;
    if (g == 1) it.next();
    for (int i = g; i < n_my; i += 2, it.next(), it.next()) {
      const int ab = g;
      unsigned char* wbuf = wb + ab * WBUF;
      const uint32_t tb = tm + ab * ACC_STRIDE + lane_sel;
      const int slot = it.slot;
      const bool warm = it.warm();
      const int tt = it.k * 128 + r;
      // this row's 16 dz values (skip-path gradient, plane l) come straight from global memory: the producer prefetched
      // the tile into the L2 when it issued the stage's loads, and the L2 round trip hides behind the wait for the MMAs.
      // (Requested a whole tile ahead from DRAM instead, the loads made the arithmetic itself ~1000 cycles slower: a
      // pending global load shares its scoreboard with the short-latency operations of the gate arithmetic.)
      uint4 dzr[2];
      dzr[0] = dzr[1] = make_uint4(0u, 0u, 0u, 0u);
      if (tt < a.T) {
        const uint4* row = reinterpret_cast<const uint4*>(a.dz + ((size_t)slot * a.T + tt) * D + 16 * half);
        dzr[0] = ldg_nc_v4_pinned(row);
        dzr[1] = ldg_nc_v4_pinned(row + 1);
      }
      int gid = 0;  // this row's voice id (global conditioning)
      if constexpr (GC) gid = tt < a.T ? min(max(__ldg(a.ids + (size_t)slot * a.T + tt), 0), a.C1 - 1) : 0;
      uint4 lcs[2], lcg[2];
      uint4* lcrow = nullptr;
      if constexpr (LC) {
        lcs[0] = lcs[1] = lcg[0] = lcg[1] = make_uint4(0u, 0u, 0u, 0u);
        if (tt < a.T) {
          const uint4* crow = reinterpret_cast<const uint4*>(a.cond + ((size_t)slot * a.T + tt) * (2 * D) + 16 * half);
          lcs[0] = crow[0]; lcs[1] = crow[1];
          lcg[0] = crow[D / 8]; lcg[1] = crow[D / 8 + 1];
          // dv goes to the gradient plane (a warm-up tile is another CTA's: only its P0 is wanted here)
          if (!warm) lcrow = reinterpret_cast<uint4*>(a.dcond + ((size_t)slot * a.T + tt) * (2 * D) + 16 * half);
        }
      }
      tr.ev(5, i);
      mbar_wait(&v_full[ab], (uint32_t)(i >> 1) & 1u);
      tr.ev(6, i);
      tc_fence_after_sync();
      if (i >= 2) mbar_wait(&p_full[ab], (uint32_t)((i - 2) >> 1) & 1u);  // the MMAs of tile i-2 have finished reading wb[ab]
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        uint32_t pz[4], pvs[4], pvg[4];
        const int c0 = 16 * half + 8 * p;
        uint32_t vs[8], vg[8], vd[8];
        tmem_ld_32x32b_x8(tb + ACC_V + c0, vs);
        tmem_ld_32x32b_x8(tb + ACC_V + 32 + c0, vg);
        if (a.has_next) {
          tmem_ld_32x32b_x8(tb + ACC_D + c0, vd);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) vd[j] = 0u;
        }
        const uint32_t dzs[4] = {dzr[p].x, dzr[p].y, dzr[p].z, dzr[p].w};
        tmem_ld_wait();
        if (p == 1) {
          tc_fence_before_sync();
          mbar_arrive(&acc1_free[ab]);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          // per-row conditioning on top of the biases (which are in the accumulator already)
          float4 b_s = make_float4(0.f, 0.f, 0.f, 0.f), b_g = make_float4(0.f, 0.f, 0.f, 0.f);
          if constexpr (GC) {  // the accumulator's GATE half holds 0.5 * pre-activation: the table's gate half is halved too
            const float* gct = a.gc_tbl + (size_t)gid * 2 * D + c0;
            const float4 c_s = __ldg(reinterpret_cast<const float4*>(gct) + q);
            const float4 c_g = __ldg(reinterpret_cast<const float4*>(gct + D) + q);
            b_s = c_s;
            b_g = make_float4(0.5f * c_g.x, 0.5f * c_g.y, 0.5f * c_g.z, 0.5f * c_g.w);
          }
          if constexpr (LC) {
            const uint32_t ws0 = q == 0 ? lcs[p].x : lcs[p].z, ws1 = q == 0 ? lcs[p].y : lcs[p].w;
            const uint32_t wg0 = q == 0 ? lcg[p].x : lcg[p].z, wg1 = q == 0 ? lcg[p].y : lcg[p].w;
            add_bf16x4(b_s, ws0, ws1, 1.f);
            add_bf16x4(b_g, wg0, wg1, 0.5f);
          }
          const float bsv[4] = {b_s.x, b_s.y, b_s.z, b_s.w}, bgv[4] = {b_g.x, b_g.y, b_g.z, b_g.w};
          float zz[4], ds[4], dg[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int d = 4 * q + k;
            float as = __uint_as_float(vs[d]), ag = __uint_as_float(vg[d]);   // ag: 0.5 * gate pre-activation
            if constexpr (GC || LC) { as += bsv[k]; ag += bgv[k]; }
            const float th = tanh_fast(as);
            const float u = tanh_fast(ag);
            const float sg = fmaf(0.5f, u, 0.5f);
            const uint32_t dw = dzs[d >> 1];
            const float dz = ((d & 1) == 0 ? __uint_as_float(dw << 16) : __uint_as_float(dw & 0xffff0000u)) + __uint_as_float(vd[d]);
            zz[k] = th * sg;
            ds[k] = (dz * sg) * fmaf(-th, th, 1.f);
            dg[k] = (dz * th) * fmaf(-0.5f * u, u, 0.5f);   // = 2 * dz th sg (1 - sg): pairs with the 0.5-scaled GATE filter
          }
          pz[2 * q] = pack2(zz[0], zz[1]);  pz[2 * q + 1] = pack2(zz[2], zz[3]);
          pvs[2 * q] = pack2(ds[0], ds[1]); pvs[2 * q + 1] = pack2(ds[2], ds[3]);
          pvg[2 * q] = pack2(dg[0], dg[1]); pvg[2 * q + 1] = pack2(dg[2], dg[3]);
          if constexpr (LC) {  // gradient wrt the conditioning plane: the true dv_g is half of what pairs with the 0.5-scaled filter
            if (q == 0) { lcg[p].x = pack2(0.5f * dg[0], 0.5f * dg[1]); lcg[p].y = pack2(0.5f * dg[2], 0.5f * dg[3]); }
            else        { lcg[p].z = pack2(0.5f * dg[0], 0.5f * dg[1]); lcg[p].w = pack2(0.5f * dg[2], 0.5f * dg[3]); }
          }
          if (GC && !warm) {
            // table gradient: dTbl[id][n] += dv[n] (gate half: dv carries 2x).  A warp is 32 consecutive timesteps of
            // one slot, ids change only at file junctions: reduce over the warp when it is uniform, else per row
            const int id0 = __shfl_sync(0xffffffffu, gid, 0);
            const bool uni = __all_sync(0xffffffffu, gid == id0);
            if (uni) {
              float w[8] = {ds[0], ds[1], ds[2], ds[3], 0.5f * dg[0], 0.5f * dg[1], 0.5f * dg[2], 0.5f * dg[3]};
              const float wsum = warp_transpose_sum8(w, lane);
              const int idx = (lane >> 2) & 7;
              if ((lane & 3) == 0 && wsum != 0.f)
                atomicAdd(a.dgc_tbl + (size_t)id0 * 2 * D + (idx < 4 ? 0 : D) + c0 + 4 * q + (idx & 3), wsum);
            } else {
              float* drow = a.dgc_tbl + (size_t)gid * 2 * D + c0 + 4 * q;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if (ds[k] != 0.f) atomicAdd(drow + k, ds[k]);
                if (dg[k] != 0.f) atomicAdd(drow + D + k, 0.5f * dg[k]);
              }
            }
          }
        }
        *reinterpret_cast<uint4*>(wbuf + W_Z * PANEL + o[p]) = make_uint4(pz[0], pz[1], pz[2], pz[3]);
        *reinterpret_cast<uint4*>(wbuf + W_DVS * PANEL + o[p]) = make_uint4(pvs[0], pvs[1], pvs[2], pvs[3]);
        *reinterpret_cast<uint4*>(wbuf + W_DVG * PANEL + o[p]) = make_uint4(pvg[0], pvg[1], pvg[2], pvg[3]);
        if constexpr (LC) {
          if (lcrow != nullptr) {
            lcrow[p] = make_uint4(pvs[0], pvs[1], pvs[2], pvs[3]);
            lcrow[D / 8 + p] = lcg[p];
          }
        }
      }
      tr.ev(15, i);
      fence_proxy_async_smem();
      tr.ev(7, i);
      if (g == 0) asm volatile("bar.sync 1, 256;" ::: "memory"); else asm volatile("bar.sync 4, 256;" ::: "memory");
      tr.ev(8, i);
      if (elected) mbar_arrive(&dv_ready[ab]);
    }
    // ---- flush: weight gradients TMEM -> staging -> coalesced atomics ----
    const int cq = 2 * g + half;  // column quarter of the 64 dv columns handled by this warp in the flush
    const int c0 = 8 * cq;
    mbar_wait(&g_full, 0);
    tc_fence_after_sync();
    if (n_my > 0) mbar_wait(&stage_free[(n_my - 1) % NST], (uint32_t)((n_my - 1) / NST) & 1u);  // last TMA store drained
    asm volatile("bar.sync 3, 768;" ::: "memory");  // the row warps are out of the ring too
    if (n_my > 0) {
      // acc_w[128 x 96]: rows 0..63, columns 0..63 = conv taps (GATE columns carry 2x); rows 64..95, columns 64..95 =
      // dx_{l+1}^T . Z = RESIDUAL's gradient, transposed
      const float gsc = cq < 2 ? 1.f : 0.5f;
      if (q4 < 2) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(tm + ACC_W + lane_sel + (uint32_t)(16 * cq), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) stg[STG_WC + r * 65 + 16 * cq + j] = gsc * __uint_as_float(v[j]);
        if (cq == 0) {  // acc_b rows 0..63 (every column alike): column sums of dv = SIGNAL_BIAS | 2 x GATE_BIAS gradients
          uint32_t bv[8];
          tmem_ld_32x32b_x8(tm + ACC_B + lane_sel, bv);
          tmem_ld_wait();
          const float val = (r < 32 ? 1.f : 0.5f) * __uint_as_float(bv[0]);
          const int64_t off = r < 32 ? a.sig_b : a.gate_b;
          if (off >= 0 && val != 0.f) atomicAdd(a.grads + off + (r & 31), val);
        }
      } else if (q4 == 2 && a.has_next) {
        uint32_t w[8];
        tmem_ld_32x32b_x8(tm + ACC_W + lane_sel + (uint32_t)(2 * D + c0), w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) stg[STG_WR + (c0 + j) * 33 + (r - 64)] = __uint_as_float(w[j]);
        if (cq == 0 && a.res_b >= 0) {  // column 96 (every ones column alike): column sums of dx_{l+1} = RESIDUAL_BIAS gradient
          uint32_t bv[8];
          tmem_ld_32x32b_x8(tm + ACC_W + lane_sel + (uint32_t)(2 * D + R), bv);
          tmem_ld_wait();
          const float val = __uint_as_float(bv[0]);
          if (val != 0.f) atomicAdd(a.grads + a.res_b + (r - 64), val);
        }
      }
      asm volatile("bar.sync 5, 512;" ::: "memory");
      // dWc row m = tap*R + rr, column n: n < D -> SIGNAL[tap][rr][n], else GATE[tap][rr][n-D]
      for (int idx = et; idx < 64 * 64; idx += NE1) {
        const int m = idx >> 6, n = idx & 63;
        const float val = stg[STG_WC + m * 65 + n];
        const int tap = m >> 5, rr = m & 31;
        float* dst = a.grads + (n < D ? a.sig : a.gate) + ((size_t)tap * R + rr) * D + (n & 31);
        if (val != 0.f) atomicAdd(dst, val);
      }
      if (a.has_next) {
        for (int idx = et; idx < 32 * 32; idx += NE1) {
          const int d = idx >> 5, c = idx & 31;
          const float val = stg[STG_WR + d * 33 + c];
          if (val != 0.f) atomicAdd(a.grads + a.res + (size_t)d * R + c, val);
        }
      }
    }
  } else if (warp < 26) {
    // ===== row warps: E2.  thread <-> (row r, channels [16*rh, +16)).
    //   P0 of this tile (bf16) -> carry slot k mod n_carry; then (not for warm-up tiles)
    //   dx_l[t] = P1[t] (dx_{l+1}[t] included: two identity MMAs of queue B) + P0[t + dil] -> the stage's dead X0 panel
    //   -> TMA store.
    //   P0[t + dil] is row (r + dil) mod 128 of tile k + (r + dil) / 128: a carry slot written by this CTA one or more
    //   tiles ago (time runs backwards here), or just now; beyond the slot's last tile it is zero, which is where the
    //   gradient stops at the stage boundary (SAVE is a variable, tmodel.py:123-124).
    //   One TMEM round trip and at most one barrier per tile: in the in-kernel timeline of the first version (two round
    //   trips, two barriers) this phase took ~3300 cycles per tile and was the bound of the whole kernel.
    //   Carry hazards: tile i + 1 writes slot (k - 1) mod n_carry, which tile i never reads (n_carry = n_later + 2);
    //   tile i + 2's MMAs are not issued before every row thread has arrived on acc2_free of tile i, after its reads. =====
    const int q4 = warp & 3, rh = (warp - 18) >> 2;
    const int r = q4 * 32 + lane;
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const uint32_t sw64 = ((uint32_t)r >> 1) & 3u;
    const uint32_t oc[2] = {(uint32_t)r * 64u + ((((uint32_t)(2 * rh)) ^ sw64) << 4),
                            (uint32_t)r * 64u + ((((uint32_t)(2 * rh + 1)) ^ sw64) << 4)};
    const int NC = a.n_carry;
    const int rs = r + (a.dil & 127), dq = (a.dil >> 7) + (rs >> 7);   // P0[t + dil]: dq tiles later, row rs & 127
    const uint32_t rr = (uint32_t)(rs & 127), swr = (rr >> 1) & 3u;
    const uint32_t ocr[2] = {rr * 64u + ((((uint32_t)(2 * rh)) ^ swr) << 4), rr * 64u + ((((uint32_t)(2 * rh + 1)) ^ swr) << 4)};
    const bool own_rows = (a.dil & 127) != 0;  // some rows of this tile's outputs need this tile's own P0
    Item it = item0;
    int cslot = it.k % NC;  // carry slot of the current tile
    for (int i = 0; i < n_my; ++i) {
      const int s = i % NST, ab = i & 1;
      unsigned char* st = smem + s * STAGE;
      const uint32_t tb = tm + ab * ACC_STRIDE + lane_sel;
      const int k = it.k;
      const bool warm = it.warm();
      tr.ev(9, i);
      mbar_wait(&p_full[ab], (uint32_t)(i >> 1) & 1u);   // every MMA reading this stage has completed
      tr.ev(10, i);
      tc_fence_after_sync();
      unsigned char* cw = carry + cslot * PANEL;
      uint32_t p0[16], p1[16];
      tmem_ld_32x32b_x16(tb + ACC_P + 16 * rh, p0);
      if (!warm) tmem_ld_32x32b_x16(tb + ACC_P + R + 16 * rh, p1);
      tmem_ld_wait();
      tc_fence_before_sync();
      mbar_arrive(&acc2_free[ab]);
#pragma unroll
      for (int c = 0; c < 2; ++c)
        *reinterpret_cast<uint4*>(cw + oc[c]) =
            make_uint4(pack2(__uint_as_float(p0[8 * c]), __uint_as_float(p0[8 * c + 1])),
                       pack2(__uint_as_float(p0[8 * c + 2]), __uint_as_float(p0[8 * c + 3])),
                       pack2(__uint_as_float(p0[8 * c + 4]), __uint_as_float(p0[8 * c + 5])),
                       pack2(__uint_as_float(p0[8 * c + 6]), __uint_as_float(p0[8 * c + 7])));
      if (!warm) {
        if (own_rows) asm volatile("bar.sync 2, 256;" ::: "memory");  // this tile's P0 rows are visible to every row thread
        const int kq = k + dq;
        int rslot = cslot + dq;   // (k + dq) mod NC without a division: dq <= NC - 2
        if (rslot >= NC) rslot -= NC;
        const unsigned char* cr = carry + rslot * PANEL;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint4 n4 = make_uint4(0u, 0u, 0u, 0u);
          if (kq < TPS) n4 = *reinterpret_cast<const uint4*>(cr + ocr[c]);
          const uint32_t pw[4] = {n4.x, n4.y, n4.z, n4.w};
          uint32_t o[4];
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            o[kk] = pack2(__uint_as_float(p1[8 * c + 2 * kk]) + __uint_as_float(pw[kk] << 16),
                          __uint_as_float(p1[8 * c + 2 * kk + 1]) + __uint_as_float(pw[kk] & 0xffff0000u));
          *reinterpret_cast<uint4*>(st + P_X0 * PANEL + oc[c]) = make_uint4(o[0], o[1], o[2], o[3]);
        }
        fence_proxy_async_smem();
      }
      tr.ev(11, i);
      __syncwarp();
      if (lane == 0) mbar_arrive(&out_ready[s]);
      tr.ev(12, i);
      it.next();
      cslot = (it.k == TPS - 1) ? (TPS - 1) % NC : (cslot == 0 ? NC - 1 : cslot - 1);
    }
    asm volatile("bar.sync 3, 768;" ::: "memory");
  } else {
    // ===== warp 26: TMA-store issuer (the only thread that touches the store path; returns the stage to the ring).
    // (The RESIDUAL_BIAS column sums were tried here, on the otherwise idle lanes, from the stage's DX panel: the stage
    // was then held ~1000 cycles longer per tile, 1.47 -> 1.59 ms for the 30 layers.) =====
    if (lane == 0) {
      Item it = item0;
      for (int i = 0; i < n_my; ++i, it.next()) {
        const int s = i % NST;
        mbar_wait(&out_ready[s], (uint32_t)(i / NST) & 1u);
        if (!it.warm()) {
          tma_store_3d_hint(&map_dxo, smem + s * STAGE + P_X0 * PANEL, 0, it.k * 128, it.slot, a.pol_out);
          tma_store_commit();
          tma_store_wait_read<0>();
        }
        tr.ev(13, i);
        mbar_arrive(&stage_free[s]);
      }
      tma_store_wait_all<0>();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (kLayerTrace && a.trace != nullptr && tid == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    a.trace[32 * WN_TRACE_PER_WARP + 4 * blockIdx.x + 1] = (long long)gt;
    atomicMax(reinterpret_cast<unsigned long long*>(a.trace) + 32 * WN_TRACE_PER_WARP + 4096 + 2 * (a.seq & 63) + 1, gt);
  }
  if (warp == 1) tmem_dealloc(tm, 512);
}

// ---- host side ------------------------------------------------------------------------------------------
static int map3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1,
                 int swizzle) {
  const uint64_t dims[3] = {d0, d1, d2};
  const uint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
  const uint32_t box[3] = {b0, b1, 1};
  return make_tensor_map_bf16(out, base, 3, dims, strides, box, swizzle);
}
static int map2ds(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint32_t b0, uint32_t b1,
                  int swizzle) {
  const uint64_t dims[2] = {inner, outer};
  const uint64_t strides[1] = {inner * 2};
  const uint32_t box[2] = {b0, b1};
  return make_tensor_map_bf16(out, base, 2, dims, strides, box, swizzle);
}

// test knob: WN_PERSIST_GRID=<n> caps the grid of the persistent kernels so that small problems exercise many
// ring / phase wrap-arounds per CTA
static int persist_grid(int want) {
  const char* e = getenv("WN_PERSIST_GRID");
  if (e != nullptr && atoi(e) > 0) return std::max(1, std::min(want, atoi(e)));
  return want;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e != nullptr ? (int)strtol(e, nullptr, 0) : dflt;
}
// WN_L2HINT: bit mask of the L2 eviction hints on the layer kernels' TMA traffic.  Measured at configs[1] (ms per step
// for the 30 forward / backward layer launches): none 0.825 / 1.531; forward x' stores EVICT_LAST (4) 0.752 / 1.52 -- the
// next launch reads x' back (33.5 MB per layer: it fits in the L2 when the single-use z stash does not push it out);
// forward x[t-dil] loads and z stores EVICT_FIRST (1, 2): no further change; backward Y / P0 loads EVICT_FIRST, stores
// EVICT_LAST, x[t-dil] loads EVICT_FIRST (8, 16, 32): no change (one layer writes 67 MB and reads another 67 MB).
static int l2_hint_mask() {
  static const int m = env_int("WN_L2HINT", 4);
  return m;
}

struct LayerMaps {
  const void* ws = nullptr;
  const void* model = nullptr;
  uint64_t serial = 0;  // wn_model::serial: a new model can be allocated at a freed model's address
  int T = -1;
  std::vector<CUtensorMap> x;  // per layer: xfull_l [B][dil+T][R]
  CUtensorMap z, wc, wr, dx[2], dz, wrn;
};

bool umma_layer_supported(const wn_model* m) {
  static const bool disabled = getenv("WN_DISABLE_UMMA") != nullptr || getenv("WN_DISABLE_UMMA_LAYER") != nullptr;
  const wn_arch& a = m->a;
  return !disabled && a.n_res == 32 && a.n_dil == 32;
}

static LayerMaps* get_maps(wn_model* m, unsigned char* ws, int T, int* rc) {
  static thread_local LayerMaps cache;  // one model per process in practice; re-encoded when ws/T change
  *rc = WN_OK;
  if (cache.model == m && cache.serial == m->serial && cache.ws == ws && cache.T == T && (int)cache.x.size() == m->L)
    return &cache;
  cache.model = nullptr;
  const WorkspaceLayout& wl = m->wl;
  const wn_arch& a = m->a;
  const uint64_t B = m->n_slots, R = a.n_res, D = a.n_dil, LD = (uint64_t)m->L * D;
  cache.x.resize(m->L);
  for (int l = 0; l < m->L; ++l)
    if ((*rc = map3d(&cache.x[l], ws + wl.xfull[l], R, (uint64_t)m->layers[l].dil + T, B, (uint32_t)R, 128, (int)R * 2)))
      return nullptr;
  if ((*rc = map3d(&cache.z, ws + wl.z, LD, (uint64_t)T, B, (uint32_t)D, 128, (int)D * 2))) return nullptr;
  if ((*rc = map2ds(&cache.wc, ws + wl.wcT, R, (uint64_t)m->L * 2 * 2 * D, (uint32_t)R, (uint32_t)(2 * D), (int)R * 2))) return nullptr;
  if ((*rc = map2ds(&cache.wr, ws + wl.wrT, D, (uint64_t)m->L * R, (uint32_t)D, (uint32_t)R, (int)D * 2))) return nullptr;
  for (int i = 0; i < 2; ++i) {
    if ((*rc = map3d(&cache.dx[i], ws + wl.dx[i], R, (uint64_t)T, B, (uint32_t)R, 128, (int)R * 2))) return nullptr;
  }
  // dz: per-layer planes [L][B][T][D]; the outer TMA coordinate is l * B + slot
  if ((*rc = map3d(&cache.dz, ws + wl.dz, D, (uint64_t)T, (uint64_t)m->L * B, (uint32_t)D, 128, (int)D * 2))) return nullptr;
  if ((*rc = map2ds(&cache.wrn, ws + wl.wrN, R, (uint64_t)m->L * D, (uint32_t)R, (uint32_t)D, (int)R * 2))) return nullptr;
  cache.ws = ws;
  cache.T = T;
  cache.model = m;
  cache.serial = m->serial;
  return &cache;
}

int launch_prep_layer_umma(wn_model* m, const float* d_params, unsigned char* ws, cudaStream_t st) {
  const WorkspaceLayout& wl = m->wl;
  k_prep_layer_weights<<<m->L, 256, 0, st>>>(d_params, m->d_layers, m->L, m->a.n_res, m->a.n_dil,
                                             reinterpret_cast<bf16*>(ws + wl.wcT), reinterpret_cast<bf16*>(ws + wl.wrT),
                                             reinterpret_cast<bf16*>(ws + wl.wrN));
  WN_LAUNCH_CHECK();
  return WN_OK;
}

int launch_layer_fwd_umma(wn_model* m, const float* d_params, unsigned char* ws, const int32_t* d_ids, int T, int l,
                          cudaStream_t st) {
  int rc;
  LayerMaps* mp = get_maps(m, ws, T, &rc);
  if (!mp) return rc;
  const WorkspaceLayout& wl = m->wl;
  const wn_arch& a = m->a;
  const LayerDesc& ld = m->layers[l];
  const int C1 = a.n_gc_category + 1;
  const bool last = (l + 1 == m->L);
  const CUtensorMap& mxo = last ? mp->x[l] : mp->x[l + 1];
  LayerFwdPArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.params = d_params;
  pa.sig_b = ld.sig_b; pa.gate_b = ld.gate_b; pa.res_b = ld.res_b;
  pa.T = T; pa.dil = ld.dil; pa.l = l; pa.last = last;
  pa.dil_next = last ? 0 : m->layers[l + 1].dil;
  pa.tiles_per_slot = (T + 127) / 128;
  pa.n_tiles = pa.tiles_per_slot * m->n_slots;
  pa.z_col = l * a.n_dil;
  pa.tile_ctr = reinterpret_cast<int*>(ws + wl.tile_ctr) + 4 * l;
  pa.trace = (g_trace_layer == -1 || g_trace_layer == l) ? g_trace_buf : nullptr;
  {
    const int h = l2_hint_mask();
    pa.pol_x0 = (h & 1) ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
    pa.pol_z = (h & 2) ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
    pa.pol_xout = (h & 4) ? L2_EVICT_LAST : L2_EVICT_NORMAL;
  }
  // stages 5 x 16 KB | z tiles 2 x 8 KB | wc 2 x 4 KB | wr 2 KB
  const size_t smem = 5 * 2 * 8192 + 2 * 8192 + 2 * 4096 + 2048 + 1024;
  const int nblk = persist_grid(std::max(1, std::min(pa.n_tiles, 2 * m->sm_count)));
  ProfScope ps(PROF_LAYER_FWD, st);
  const bool gc = a.n_gc_embed > 0, lc = a.n_lc_out > 0;
  if (gc) {
    pa.gc_tbl = reinterpret_cast<const float*>(ws + wl.gc_tbl) + (size_t)l * C1 * 2 * a.n_dil;
    pa.ids = d_ids;
    pa.C1 = C1;
  }
  if (lc) pa.cond = reinterpret_cast<const bf16*>(ws + wl.cond) + (size_t)l * m->n_slots * T * 2 * a.n_dil;
#define WN_FWD(GC_, LC_)                                                                                                  \
  {                                                                                                                       \
    WN_CUDA_CHECK(cudaFuncSetAttribute(k_layer_fwd_p_umma<32, 32, GC_, LC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                       (int)smem));                                                                       \
    WN_CUDA_CHECK(launch_pdl(k_layer_fwd_p_umma<32, 32, GC_, LC_>, nblk, 608, smem, st, mp->x[l], mxo, mp->z, mp->wc,     \
                             mp->wr, pa));                                                                                \
  }
  if (gc && lc) WN_FWD(true, true)
  else if (gc) WN_FWD(true, false)
  else if (lc) WN_FWD(false, true)
  else WN_FWD(false, false)
#undef WN_FWD
  WN_LAUNCH_CHECK();
  return WN_OK;
}

bool umma_bwd_fused_supported(const wn_model* m) { return umma_layer_supported(m); }

// whole backward of layer l: reads dx_{l+1} from dxbuf[(l+1) & 1], writes dx_l to dxbuf[l & 1]
int launch_layer_bwd_fused_umma(wn_model* m, const float* d_params, unsigned char* ws, const int32_t* d_ids, int T, int l,
                                float* d_grads, cudaStream_t st) {
  int rc;
  LayerMaps* mp = get_maps(m, ws, T, &rc);
  if (!mp) return rc;
  const LayerDesc& ld = m->layers[l];
  LayerBwdFusedArgs ga;
  memset(&ga, 0, sizeof(ga));
  ga.params = d_params;
  ga.grads = d_grads;
  ga.sig = ld.sig; ga.gate = ld.gate; ga.res = ld.res;
  ga.sig_b = ld.sig_b; ga.gate_b = ld.gate_b; ga.res_b = ld.res_b;
  ga.dil = ld.dil; ga.l = l;
  ga.has_next = (l + 1 < m->L);
  ga.tiles_per_slot = (T + 127) / 128;
  ga.n_tiles = ga.tiles_per_slot * m->n_slots;
  ga.z_plane0 = l * m->n_slots;
  ga.trace = (g_trace_layer == -1 || g_trace_layer == l) ? g_trace_buf : nullptr;
  static int trace_seq = 0;
  if (ga.trace != nullptr) ga.seq = trace_seq++;
  // ring n_stages x 24 KB | work buffers 2 x 32 KB | wc 2 x 4 KB | RESIDUAL 2 KB | bias tile 2 KB | [0 | I] 4 KB | carry n_carry x 8 KB
  ga.n_later = (ld.dil + 127) / 128;
  ga.n_carry = ga.n_later + 2;  // one spare: the next tile's P0 never lands in a slot the current tile still reads
  const int fixed = 2 * 4 * 8192 + 2 * 4096 + 2048 + 2048 + 4096 + ga.n_carry * 8192 + 1024;
  ga.n_stages = std::min(6, (232448 /* 227 KB per CTA on sm_100 */ - 2048 - fixed) / (3 * 8192));
  if (ga.n_stages < 2) {
    set_error("layer backward: dilation %d needs %d carry tiles, more than shared memory holds", ld.dil, ga.n_carry);
    return WN_ERR_UNSUPPORTED;
  }
  const size_t smem = (size_t)ga.n_stages * 3 * 8192 + fixed;
  const int grid = persist_grid(std::max(1, std::min(ga.n_tiles, m->sm_count)));
  const int nx = (l + 1) & 1, cu = l & 1;
  ProfScope ps(PROF_LAYER_BWD_A, st);
  const bool gc = m->a.n_gc_embed > 0, lc = m->a.n_lc_out > 0;
  ga.T = T;
  {
    const int h = l2_hint_mask();
    ga.pol_in = (h & 8) ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
    ga.pol_out = (h & 16) ? L2_EVICT_LAST : L2_EVICT_NORMAL;
    ga.pol_x0 = (h & 32) ? L2_EVICT_FIRST : L2_EVICT_NORMAL;
  }
  ga.dz = reinterpret_cast<const bf16*>(ws + m->wl.dz) + (size_t)l * m->n_slots * T * m->a.n_dil;
  if (gc) {
    const int C1 = m->a.n_gc_category + 1;
    ga.gc_tbl = reinterpret_cast<const float*>(ws + m->wl.gc_tbl) + (size_t)l * C1 * 2 * m->a.n_dil;
    ga.dgc_tbl = reinterpret_cast<float*>(ws + m->wl.dgc_tbl) + (size_t)l * C1 * 2 * m->a.n_dil;
    ga.ids = d_ids;
    ga.C1 = C1;
  }
  if (lc) {
    ga.cond = reinterpret_cast<const bf16*>(ws + m->wl.cond) + (size_t)l * m->n_slots * T * 2 * m->a.n_dil;
    ga.dcond = reinterpret_cast<bf16*>(ws + m->wl.dcond) + (size_t)l * m->n_slots * T * 2 * m->a.n_dil;
  }
#define WN_BWD(GC_, LC_)                                                                                                  \
  {                                                                                                                       \
    WN_CUDA_CHECK(cudaFuncSetAttribute(k_layer_bwd_fused_umma<32, 32, GC_, LC_>,                                          \
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                          \
    WN_CUDA_CHECK(launch_pdl(k_layer_bwd_fused_umma<32, 32, GC_, LC_>, grid, 896, smem, st, mp->x[l], mp->dz,             \
                             mp->dx[nx], mp->dx[cu], mp->wc, mp->wrn, ga));                                               \
  }
  if (gc && lc) WN_BWD(true, true)
  else if (gc) WN_BWD(true, false)
  else if (lc) WN_BWD(false, true)
  else WN_BWD(false, false)
#undef WN_BWD
  WN_LAUNCH_CHECK();
  return WN_OK;
}

}  // namespace wn
