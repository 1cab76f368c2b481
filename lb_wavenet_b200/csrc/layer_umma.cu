// Per-layer kernels, tcgen05 generation (R = D in {32, 64}: one activation row is one swizzle span).
//
// k_layer_fwd_umma  (reference tmodel.py:117-168 _dilated_conv, :171-184 _chan_reduce, :325 residual add)
//   one CTA = one 128-timestep tile of one slot; activations live in the "prefix" layout
//   xfull_l [slot][dil_l + T][R] (rows [0, dil) = saved D-separation state, tmodel.py:127), so both conv
//   taps are plain TMA boxes of the same 3-D tensor at row t0 (x[t-dil]) and t0 + dil (x[t]):
//     acc_v[128 x 2D] = x[t-dil] . W[0] + x[t] . W[1]      (SIGNAL | GATE side by side in N)
//     z = bf16(tanh(v_s + b_s [+gc]) * sigmoid(v_g + b_g [+gc]))  -> smem (A of the next MMA) + TMA store
//     acc_r[128 x R]  = z . RESIDUAL ;  x' = bf16(x[t] + acc_r + b_r)  -> TMA store into xfull_{l+1}
// k_layer_bwd_dx_umma: dx_l[t] = dx_{l+1}[t] + dv[t] . W[1]^T + dv[t+dil] . W[0]^T  (rows t+dil >= T are
//   zero-filled by TMA: the gradient stops at the stage boundary, SAVE being a variable not a graph tensor)
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

// ---- weight preparation: K-major B operands -------------------------------------------------------
// wcT[l][tap][n][r] = (n < D ? SIGNAL : GATE)[tap][r][n % D]      ([2D rows][R], one block per tap)
// wrT[l][r][d]      = RESIDUAL[d][r]                               ([R rows][D])
// wdT[l][tap][r][n] = (n < D ? SIGNAL : GATE)[tap][r][n % D]      ([R rows][2D]: data-gradient B operand)
// wr [l][d][r]      = RESIDUAL[d][r] (bf16 copy, [D rows][R]: B operand of dz = dx' . RESIDUAL^T)
__global__ void k_prep_layer_weights(const float* __restrict__ p, const LayerDesc* __restrict__ layers, int L, int R,
                                     int D, bf16* __restrict__ wcT, bf16* __restrict__ wrT, bf16* __restrict__ wdT,
                                     bf16* __restrict__ wr) {
  const int l = blockIdx.x;
  const LayerDesc ld = layers[l];
  const int n_wc = 2 * 2 * D * R, n_wr = R * D;
  for (int i = threadIdx.x; i < n_wc; i += blockDim.x) {
    const int tap = i / (2 * D * R), rem = i % (2 * D * R);
    const int n = rem / R, r = rem % R;
    const float v = p[(n < D ? ld.sig : ld.gate) + ((int64_t)tap * R + r) * D + (n % D)];
    wcT[(int64_t)l * n_wc + i] = f2bf(v);
    wdT[(int64_t)l * n_wc + ((int64_t)tap * R + r) * 2 * D + n] = f2bf(v);
  }
  for (int i = threadIdx.x; i < n_wr; i += blockDim.x) {
    const int r = i / D, d = i % D;
    const float v = p[ld.res + (int64_t)d * R + r];
    wrT[(int64_t)l * n_wr + i] = f2bf(v);
    wr[(int64_t)l * n_wr + (int64_t)d * R + r] = f2bf(v);
  }
}

struct LayerFwdUmmaArgs {
  const float* params;
  int64_t sig_b, gate_b, res_b;  // -1: no bias
  const float* gc_tbl;            // this layer's [C1][2D] table or nullptr
  const int32_t* ids;
  int T, dil, dil_next, l, C1, last;
};

template <int R, int D>
__global__ void __launch_bounds__(128)
k_layer_fwd_umma(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_xout,
                 const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_wc,
                 const __grid_constant__ CUtensorMap map_wr, LayerFwdUmmaArgs a) {
  constexpr int XB = R * 2, ZB = D * 2;            // bytes per row == swizzle span
  constexpr int X_TILE = 128 * XB, Z_TILE = 128 * ZB;
  constexpr int WC_TILE = 2 * D * XB, WR_TILE = R * ZB;
  static_assert(XB <= 128 && ZB <= 128, "one row must fit one swizzle span");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* x0 = smem;
  unsigned char* x1 = x0 + X_TILE;
  unsigned char* ztile = x1 + X_TILE;
  unsigned char* otile = x0;  // x[t-dil] is only read by the first MMA, complete before the output tile is written
  unsigned char* wc0 = ztile + Z_TILE;
  unsigned char* wc1 = wc0 + WC_TILE;
  unsigned char* wr = wc1 + WC_TILE;
  __shared__ __align__(8) uint64_t bar_in, bar_v, bar_r;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.y, t0 = blockIdx.x * 128;
  constexpr uint32_t NCOL = 2 * D <= 64 ? 64 : (2 * D <= 128 ? 128 : 256);  // acc_r reuses acc_v's columns

  if (tid == 0) {
    mbar_init(&bar_in, 1);
    mbar_init(&bar_v, 1);
    mbar_init(&bar_r, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, NCOL);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t acc_v = tmem_base_s, acc_r = tmem_base_s;  // second MMA is issued after every thread drained acc_v

  if (tid == 0) {
    mbar_expect_tx(&bar_in, (uint32_t)(2 * X_TILE + 2 * WC_TILE + (a.last ? 0 : WR_TILE)));
    tma_load_3d(x0, &map_x, &bar_in, 0, t0, b);                 // x[t - dil]
    tma_load_3d(x1, &map_x, &bar_in, 0, t0 + a.dil, b);         // x[t]
    tma_load_2d(wc0, &map_wc, &bar_in, 0, (a.l * 2 + 0) * 2 * D);
    tma_load_2d(wc1, &map_wc, &bar_in, 0, (a.l * 2 + 1) * 2 * D);
    if (!a.last) tma_load_2d(wr, &map_wr, &bar_in, 0, a.l * R);
    mbar_wait(&bar_in, 0);
    tc_fence_after_sync();
    const uint32_t idesc = make_idesc_bf16(128, 2 * D);
#pragma unroll
    for (int k = 0; k < R / 16; ++k)
      mma_bf16_ss(acc_v, make_kmajor_desc(smem_u32(x0), XB, k * 32), make_kmajor_desc(smem_u32(wc0), XB, k * 32), idesc, k != 0);
#pragma unroll
    for (int k = 0; k < R / 16; ++k)
      mma_bf16_ss(acc_v, make_kmajor_desc(smem_u32(x1), XB, k * 32), make_kmajor_desc(smem_u32(wc1), XB, k * 32), idesc, true);
    mma_commit(&bar_v);
  }
  // ---- gate: thread <-> row ----
  const int r = tid, t = t0 + r;
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
  const float* gct = nullptr;
  if (a.gc_tbl != nullptr && t < a.T) {
    int id = a.ids[(size_t)b * a.T + t];
    id = min(max(id, 0), a.C1 - 1);
    gct = a.gc_tbl + (size_t)id * 2 * D;
  }
  mbar_wait(&bar_v, 0);
  tc_fence_after_sync();
  {
    uint32_t vs[32], vg[32], pk[16];
#pragma unroll
    for (int c0 = 0; c0 < D; c0 += 32) {
      tmem_ld_32x32b_x32(acc_v + lane_sel + (uint32_t)c0, vs);
      tmem_ld_32x32b_x32(acc_v + lane_sel + (uint32_t)(D + c0), vg);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float z[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int d = c0 + 2 * j + e;
          float s = __uint_as_float(vs[2 * j + e]), g = __uint_as_float(vg[2 * j + e]);
          if (a.sig_b >= 0) {
            s += __ldg(a.params + a.sig_b + d);
            g += __ldg(a.params + a.gate_b + d);
          }
          if (gct != nullptr) {
            s += __ldg(gct + d);
            g += __ldg(gct + D + d);
          }
          z[e] = tanh_fast(s) * sigmoid_fast(g);
        }
        pk[j] = pack2(z[0], z[1]);
      }
      row_store<ZB, 64>(ztile, r, c0 * 2, pk);
    }
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after_sync();
    if (!a.last) {
      const uint32_t idesc = make_idesc_bf16(128, R);
#pragma unroll
      for (int k = 0; k < D / 16; ++k)
        mma_bf16_ss(acc_r, make_kmajor_desc(smem_u32(ztile), ZB, k * 32), make_kmajor_desc(smem_u32(wr), ZB, k * 32), idesc, k != 0);
      mma_commit(&bar_r);
    }
    tma_store_3d(&map_z, ztile, a.l * D, t0, b);
    tma_store_commit();
  }
  if (!a.last) {
    mbar_wait(&bar_in, 0);  // the TMA-written x[t] tile is visible to this thread's generic loads
    mbar_wait(&bar_r, 0);
    tc_fence_after_sync();
    uint32_t vr[32], xin[16], pk[16];
#pragma unroll
    for (int c0 = 0; c0 < R; c0 += 32) {
      tmem_ld_32x32b_x32(acc_r + lane_sel + (uint32_t)c0, vr);
      row_load<XB, 64>(x1, r, c0 * 2, xin);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float o0 = __uint_as_float(vr[2 * j]) + __uint_as_float(xin[j] << 16);
        float o1 = __uint_as_float(vr[2 * j + 1]) + __uint_as_float(xin[j] & 0xffff0000u);
        if (a.res_b >= 0) {
          o0 += __ldg(a.params + a.res_b + c0 + 2 * j);
          o1 += __ldg(a.params + a.res_b + c0 + 2 * j + 1);
        }
        pk[j] = pack2(o0, o1);
      }
      row_store<XB, 64>(otile, r, c0 * 2, pk);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tma_store_3d(&map_xout, otile, 0, a.dil_next + t0, b);
      tma_store_commit();
    }
  }
  if (tid == 0) tma_store_wait_all<0>();
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, NCOL);
}

// ---- data gradient ------------------------------------------------------------------------------------
struct LayerDxUmmaArgs {
  int dil, l, has_next;
};

template <int R, int D>
__global__ void __launch_bounds__(128)
k_layer_bwd_dx_umma(const __grid_constant__ CUtensorMap map_dv, const __grid_constant__ CUtensorMap map_wd,
                    const __grid_constant__ CUtensorMap map_dxn, const __grid_constant__ CUtensorMap map_dxo,
                    LayerDxUmmaArgs a) {
  constexpr int XB = R * 2, VB = 2 * D * 2;  // dv rows: 2D bf16
  constexpr int X_TILE = 128 * XB, V_TILE = 128 * VB, WD_TILE = R * VB;
  static_assert(VB <= 128, "dv row must fit one swizzle span");
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* dva = smem;
  unsigned char* dvb = dva + V_TILE;
  unsigned char* wd0 = dvb + V_TILE;
  unsigned char* wd1 = wd0 + WD_TILE;
  unsigned char* dxn = wd1 + WD_TILE;
  unsigned char* otile = dva;  // dv[t] is only read by the MMAs, complete before the output tile is written
  __shared__ __align__(8) uint64_t bar_in, bar_acc;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.y, t0 = blockIdx.x * 128;
  constexpr uint32_t NCOL = R <= 32 ? 32 : (R <= 64 ? 64 : 128);
  if (tid == 0) {
    mbar_init(&bar_in, 1);
    mbar_init(&bar_acc, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, NCOL);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t acc = tmem_base_s;
  if (tid == 0) {
    mbar_expect_tx(&bar_in, (uint32_t)(2 * V_TILE + 2 * WD_TILE + (a.has_next ? X_TILE : 0)));
    tma_load_3d(dva, &map_dv, &bar_in, 0, t0, b);
    tma_load_3d(dvb, &map_dv, &bar_in, 0, t0 + a.dil, b);  // rows >= T: zero fill == truncated gradient
    tma_load_2d(wd0, &map_wd, &bar_in, 0, (a.l * 2 + 0) * R);
    tma_load_2d(wd1, &map_wd, &bar_in, 0, (a.l * 2 + 1) * R);
    if (a.has_next) tma_load_3d(dxn, &map_dxn, &bar_in, 0, t0, b);
    mbar_wait(&bar_in, 0);
    tc_fence_after_sync();
    const uint32_t idesc = make_idesc_bf16(128, R);
#pragma unroll
    for (int k = 0; k < 2 * D / 16; ++k)  // dv[t] . W[1]^T
      mma_bf16_ss(acc, make_kmajor_desc(smem_u32(dva), VB, k * 32), make_kmajor_desc(smem_u32(wd1), VB, k * 32), idesc, k != 0);
#pragma unroll
    for (int k = 0; k < 2 * D / 16; ++k)  // dv[t+dil] . W[0]^T
      mma_bf16_ss(acc, make_kmajor_desc(smem_u32(dvb), VB, k * 32), make_kmajor_desc(smem_u32(wd0), VB, k * 32), idesc, true);
    mma_commit(&bar_acc);
  }
  const int r = tid;
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
  mbar_wait(&bar_acc, 0);
  tc_fence_after_sync();
  if (a.has_next) mbar_wait(&bar_in, 0);  // dx' tile visible to every thread
  uint32_t vr[32], xin[16], pk[16];
#pragma unroll
  for (int c0 = 0; c0 < R; c0 += 32) {
    tmem_ld_32x32b_x32(acc + lane_sel + (uint32_t)c0, vr);
    if (a.has_next) {
      row_load<XB, 64>(dxn, r, c0 * 2, xin);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) xin[j] = 0u;
    }
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      pk[j] = pack2(__uint_as_float(vr[2 * j]) + __uint_as_float(xin[j] << 16),
                    __uint_as_float(vr[2 * j + 1]) + __uint_as_float(xin[j] & 0xffff0000u));
    row_store<XB, 64>(otile, r, c0 * 2, pk);
  }
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  if (tid == 0) {
    tma_store_3d(&map_dxo, otile, 0, t0, b);
    tma_store_commit();
    tma_store_wait_all<0>();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, NCOL);
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// =====================================================================================================
// k_layer_bwd_gate_umma: persistent gate-backward kernel with in-TMEM weight-gradient accumulation.
//   per 128-timestep tile (one slot):
//     acc_v = x[t-dil] . W[0] + x[t] . W[1]                  (recomputed pre-activations, SIGNAL | GATE)
//     acc_d = dx_{l+1}[t] . RESIDUAL^T                       (residual part of dz)
//     th = tanh(v_s + b), sg = sigmoid(v_g + b), z = th * sg (bf16, the tile the forward stored)
//     dz = dz_skip (from the post-net backward) + acc_d
//     dv = [dz * sg * (1 - th^2) | dz * th * sg * (1 - sg)]   -> bf16 tile -> TMA store (data-gradient kernel)
//   weight gradients, accumulated across ALL tiles of the CTA in tensor memory (MN-major operands are the
//   very same shared-memory tiles, re-described):
//     acc_wc[0:64 , 0:64] += [x[t-dil] | x[t]]^T . dv         (SIGNAL / GATE taps 0 and 1)
//     acc_wr[64:96, 0:32] += z^T . dx_{l+1}                    (RESIDUAL)
//   bias gradients = column sums of the dv / dx_{l+1} tiles, accumulated in registers across tiles.
//   One flush (coalesced fp32 atomics) per CTA at the end.
// =====================================================================================================
struct LayerGateUmmaArgs {
  const float* params;
  float* grads;
  int64_t sig, gate, res, sig_b, gate_b, res_b;
  int T, dil, l, has_next, n_tiles, tiles_per_slot;
};

template <int R, int D>
__global__ void __launch_bounds__(320, 1)
k_layer_bwd_gate_umma(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dz,
                      const __grid_constant__ CUtensorMap map_dxn, const __grid_constant__ CUtensorMap map_dv,
                      const __grid_constant__ CUtensorMap map_wc, const __grid_constant__ CUtensorMap map_wrn,
                      LayerGateUmmaArgs a) {
  static_assert(R == 32 && D == 32, "tile bookkeeping below assumes 64-byte activation rows");
  constexpr int XB = 64, VB = 128;
  constexpr int PANEL = 128 * XB;                 // 8 KB: one [128 x 32] bf16 tile
  // stage: x0 | x1 | z | ones | dz | dxn | dv(16 KB).  Panels 0..3 are the MN-major A operand of the weight-gradient
  // MMAs (M = 128: rows 0..63 conv taps, 64..95 z, 96..127 constant one -> bias gradients for free)
  constexpr int STAGE = 6 * PANEL + 128 * VB;
  constexpr int NST = 3;  // two tiles of TMA prefetch distance (the kernel is load-latency bound otherwise)
  constexpr int NEPI = 256;                       // 8 epilogue warps: (TMEM lane quarter) x (channel half)
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* wc0 = smem + NST * STAGE;        // [2D rows][R]   4 KB
  unsigned char* wc1 = wc0 + 2 * D * XB;
  unsigned char* wrn = wc1 + 2 * D * XB;          // [D rows][R]    2 KB
  float* stg = reinterpret_cast<float*>(smem);    // end-of-kernel staging (aliases the stages)
  __shared__ __align__(8) uint64_t w_full, in_full[NST], stage_free[NST], v_full[2], acc_free[2], dv_ready[NST], g_full;
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_sm[64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_my = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (tid == 0) {
    mbar_init(&w_full, 1);
    mbar_init(&g_full, 1);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&in_full[i], 1);
      mbar_init(&stage_free[i], 2);
      mbar_init(&dv_ready[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&acc_free[i], NEPI);
    }
    fence_mbar_init();
  }
  if (tid < 64) bias_sm[tid] = tid < 32 ? (a.sig_b >= 0 ? a.params[a.sig_b + tid] : 0.f)
                                        : (a.gate_b >= 0 ? a.params[a.gate_b + tid - 32] : 0.f);
  // constant-one panels (bf16 1.0 = 0x3f80); the swizzle permutes equal values, so a plain fill is exact
  for (int i = tid; i < NST * PANEL / 16; i += blockDim.x) {
    const int s = i / (PANEL / 16), o = i % (PANEL / 16);
    *reinterpret_cast<uint4*>(smem + s * STAGE + 3 * PANEL + o * 16) = make_uint4(0x3f803f80u, 0x3f803f80u, 0x3f803f80u, 0x3f803f80u);
  }
  fence_proxy_async_smem();
  if (warp == 1) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = tmem_base_s;
  // TMEM columns: per-tile buffers ab in {0,1}: acc_v at ab*128 (64 cols), acc_d at ab*128 + 64 (32 cols);
  // persistent: acc_wc at 256 (64 cols), acc_wr at 320 (32 cols)
  const uint32_t acc_wc = tm + 256, acc_wr = tm + 320;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&w_full, (uint32_t)(2 * 2 * D * XB + D * XB));
      tma_load_2d(wc0, &map_wc, &w_full, 0, (a.l * 2 + 0) * 2 * D);
      tma_load_2d(wc1, &map_wc, &w_full, 0, (a.l * 2 + 1) * 2 * D);
      tma_load_2d(wrn, &map_wrn, &w_full, 0, a.l * D);
      for (int i = 0; i < n_my; ++i) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int b = tile / a.tiles_per_slot, t0 = (tile % a.tiles_per_slot) * 128;
        const int s = i % NST;
        unsigned char* st = smem + s * STAGE;
        mbar_wait(&stage_free[s], ((uint32_t)(i / NST) & 1u) ^ 1u);
        mbar_expect_tx(&in_full[s], (uint32_t)((a.has_next ? 4 : 3) * PANEL));
        tma_load_3d(st, &map_x, &in_full[s], 0, t0, b);
        tma_load_3d(st + PANEL, &map_x, &in_full[s], 0, t0 + a.dil, b);
        tma_load_3d(st + 4 * PANEL, &map_dz, &in_full[s], a.l * D, t0, b);
        if (a.has_next) tma_load_3d(st + 5 * PANEL, &map_dxn, &in_full[s], 0, t0, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(&w_full, 0);
      const uint32_t idv = make_idesc_bf16(128, 2 * D), idd = make_idesc_bf16(128, D);
      const uint32_t idwc = make_idesc_bf16(128, 2 * D, true, true), idwr = make_idesc_bf16(128, R, true, true);
      auto issue_wgrad = [&](int j) {  // weight gradients of tile j (its dv / z tiles are in shared memory)
        const int s = j % NST;
        const uint32_t st = smem_u32(smem + s * STAGE);
        mbar_wait(&dv_ready[s], (uint32_t)(j / NST) & 1u);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 8; ++k)  // K = 128 timesteps, 16 per instruction
          mma_bf16_ss(acc_wc, make_mnmajor_desc(st + k * 16 * XB, XB, PANEL),
                      make_mnmajor_desc(st + 6 * PANEL + k * 16 * VB, VB, 0), idwc, (j | k) != 0);
        if (a.has_next) {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            mma_bf16_ss(acc_wr, make_mnmajor_desc(st + k * 16 * XB, XB, PANEL),
                        make_mnmajor_desc(st + 5 * PANEL + k * 16 * XB, XB, 0), idwr, (j | k) != 0);
        }
        mma_commit(&stage_free[s]);
      };
      for (int i = 0; i < n_my; ++i) {
        const int s = i % NST, ab = i & 1;
        const uint32_t st = smem_u32(smem + s * STAGE);
        mbar_wait(&in_full[s], (uint32_t)(i / NST) & 1u);
        mbar_wait(&acc_free[ab], ((uint32_t)(i >> 1) & 1u) ^ 1u);
        tc_fence_after_sync();
        const uint32_t av = tm + ab * 128, ad = av + 64;
#pragma unroll
        for (int k = 0; k < R / 16; ++k)
          mma_bf16_ss(av, make_kmajor_desc(st, XB, k * 32), make_kmajor_desc(smem_u32(wc0), XB, k * 32), idv, k != 0);
#pragma unroll
        for (int k = 0; k < R / 16; ++k)
          mma_bf16_ss(av, make_kmajor_desc(st + PANEL, XB, k * 32), make_kmajor_desc(smem_u32(wc1), XB, k * 32), idv, true);
        if (a.has_next) {
#pragma unroll
          for (int k = 0; k < R / 16; ++k)  // dz(res) = dx' . RESIDUAL^T : B = RESIDUAL [D rows][R]
            mma_bf16_ss(ad, make_kmajor_desc(st + 5 * PANEL, XB, k * 32), make_kmajor_desc(smem_u32(wrn), XB, k * 32), idd, k != 0);
        }
        mma_commit(&v_full[ab]);
        if (i > 0) issue_wgrad(i - 1);  // overlaps the epilogue of tile i with the tensor work of tile i-1
      }
      if (n_my > 0) issue_wgrad(n_my - 1);
      mma_commit(&g_full);
    }
  } else {
    const int e = warp - 2;                    // 0..7
    const int q4 = warp & 3, half = e >> 2;    // TMEM lane quarter, channel half [16*half, 16*half+16)
    const int r = q4 * 32 + lane;
    const int et = e * 32 + lane;              // 0..255
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const bool elected = (warp == 2 && lane == 0);
    const int c0 = 16 * half;
    auto ebar = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    for (int i = 0; i < n_my; ++i) {
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const int b = tile / a.tiles_per_slot, t0 = (tile % a.tiles_per_slot) * 128;
      const int s = i % NST, ab = i & 1;
      unsigned char* st = smem + s * STAGE;
      if (elected && i > 0) {  // release the previous tile's stage once its dv store has finished reading it
        tma_store_wait_read<0>();
        mbar_arrive(&stage_free[(i - 1) % NST]);
      }
      mbar_wait(&in_full[s], (uint32_t)(i / NST) & 1u);  // dz tile (TMA) visible to this thread
      mbar_wait(&v_full[ab], (uint32_t)(i >> 1) & 1u);
      tc_fence_after_sync();
      uint32_t vs[16], vg[16], vd[16], dzs[8], pz[8], pvs[8], pvg[8];
      tmem_ld_32x32b_x16(tm + ab * 128 + c0 + lane_sel, vs);
      tmem_ld_32x32b_x16(tm + ab * 128 + 32 + c0 + lane_sel, vg);
      if (a.has_next) tmem_ld_32x32b_x16(tm + ab * 128 + 64 + c0 + lane_sel, vd);
      row_load<XB, 32>(st + 4 * PANEL, r, 32 * half, dzs);
      tmem_ld_wait();
      tc_fence_before_sync();
      mbar_arrive(&acc_free[ab]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float zz[2], ds[2], dg[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int d = 2 * j + k;
          const float th = tanh_fast(__uint_as_float(vs[d]) + bias_sm[c0 + d]);
          const float sg = sigmoid_fast(__uint_as_float(vg[d]) + bias_sm[32 + c0 + d]);
          float dz = k == 0 ? __uint_as_float(dzs[j] << 16) : __uint_as_float(dzs[j] & 0xffff0000u);
          if (a.has_next) dz += __uint_as_float(vd[d]);
          zz[k] = th * sg;
          ds[k] = dz * sg * (1.f - th * th);
          dg[k] = dz * th * sg * (1.f - sg);
        }
        pz[j] = pack2(zz[0], zz[1]);
        pvs[j] = pack2(ds[0], ds[1]);
        pvg[j] = pack2(dg[0], dg[1]);
      }
      row_store<XB, 32>(st + 2 * PANEL, r, 32 * half, pz);            // z tile (A panel 2 of the weight-gradient MMA)
      row_store<VB, 32>(st + 6 * PANEL, r, 32 * half, pvs);           // dv tile: signal half | gate half
      row_store<VB, 32>(st + 6 * PANEL, r, 64 + 32 * half, pvg);
      fence_proxy_async_smem();
      ebar();
      if (elected) {
        tma_store_3d(&map_dv, st + 6 * PANEL, 0, t0, b);
        tma_store_commit();
        mbar_arrive(&dv_ready[s]);
      }
    }
    if (elected && n_my > 0) {
      tma_store_wait_read<0>();
      mbar_arrive(&stage_free[(n_my - 1) % NST]);
    }
    // ---- flush: weight gradients TMEM -> staging -> coalesced atomics; bias gradients from the ones-row ----
    mbar_wait(&g_full, 0);
    tc_fence_after_sync();
    ebar();
    if (n_my > 0) {
      uint32_t v[32];
      // acc_wc[128 x 64]: rows 0..63 = conv taps (lanes 0..63), row 96 = column sums of dv (lane 96);
      // acc_wr[128 x 32]: rows 64..95 = RESIDUAL (lanes 64..95), row 96 = column sums of dx'
      if (q4 < 2) {
        tmem_ld_32x32b_x32(acc_wc + lane_sel + (uint32_t)(32 * half), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[r * 65 + 32 * half + j] = __uint_as_float(v[j]);
      } else if (q4 == 2 && a.has_next) {
        uint32_t w[16];
        tmem_ld_32x32b_x16(acc_wr + lane_sel + (uint32_t)c0, w);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) stg[64 * 65 + (r - 64) * 33 + c0 + j] = __uint_as_float(w[j]);
      } else if (q4 == 3) {
        tmem_ld_32x32b_x32(acc_wc + lane_sel + (uint32_t)(32 * half), v);
        uint32_t w[16];
        if (a.has_next) tmem_ld_32x32b_x16(acc_wr + lane_sel + (uint32_t)c0, w);
        tmem_ld_wait();
        if (lane == 0) {  // row 96
          if (a.sig_b >= 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float val = __uint_as_float(v[j]);
              if (val != 0.f) atomicAdd(a.grads + (half == 0 ? a.sig_b : a.gate_b) + j, val);
            }
          }
          if (a.has_next && a.res_b >= 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float val = __uint_as_float(w[j]);
              if (val != 0.f) atomicAdd(a.grads + a.res_b + c0 + j, val);
            }
          }
        }
      }
      ebar();
      // dWc row m = tap*R + rr, column n: n < D -> SIGNAL[tap][rr][n], else GATE[tap][rr][n-D]
      for (int idx = et; idx < 64 * 64; idx += NEPI) {
        const int m = idx >> 6, n = idx & 63;
        const float val = stg[m * 65 + n];
        const int tap = m >> 5, rr = m & 31;
        float* dst = a.grads + (n < D ? a.sig : a.gate) + ((size_t)tap * R + rr) * D + (n & 31);
        if (val != 0.f) atomicAdd(dst, val);
      }
      if (a.has_next) {
        for (int idx = et; idx < 32 * 32; idx += NEPI) {
          const int d = idx >> 5, c = idx & 31;
          const float val = stg[64 * 65 + d * 33 + c];
          if (val != 0.f) atomicAdd(a.grads + a.res + (size_t)d * R + c, val);
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tm, 512);
}

// =====================================================================================================
// Persistent variants of the forward and data-gradient kernels (the per-tile work is a short latency chain
// TMA -> MMA -> epilogue -> TMA; a deep TMA ring and double-buffered TMEM keep HBM busy).
// Roles: warp 0 producer, warp 1 MMA issuer, warps 2..9 epilogue ((TMEM lane quarter) x (channel half)).
// =====================================================================================================
struct LayerFwdPArgs {
  const float* params;
  int64_t sig_b, gate_b, res_b;
  int T, dil, dil_next, l, last, n_tiles, tiles_per_slot;
};

template <int R, int D>
__global__ void __launch_bounds__(320, 2)
k_layer_fwd_p_umma(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_xout,
                   const __grid_constant__ CUtensorMap map_z, const __grid_constant__ CUtensorMap map_wc,
                   const __grid_constant__ CUtensorMap map_wr, LayerFwdPArgs a) {
  static_assert(R == 32 && D == 32, "64-byte activation rows");
  constexpr int XB = 64;
  constexpr int PANEL = 128 * XB;       // 8 KB
  constexpr int STAGE = 2 * PANEL;      // x[t-dil] | x[t]
  constexpr int NST = 4;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* zt = smem + NST * STAGE;       // [2] z tiles
  unsigned char* ot = zt + 2 * PANEL;           // [2] output tiles
  unsigned char* wc0 = ot + 2 * PANEL;
  unsigned char* wc1 = wc0 + 2 * D * XB;
  unsigned char* wr = wc1 + 2 * D * XB;
  __shared__ __align__(8) uint64_t w_full, in_full[NST], stage_free[NST], v_full[2], r_full[2], acc_free[2], z_ready[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_sm[96];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_my = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (tid == 0) {
    mbar_init(&w_full, 1);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&in_full[i], 1);
      mbar_init(&stage_free[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&r_full[i], 1);
      mbar_init(&acc_free[i], 256);
      mbar_init(&z_ready[i], 1);
    }
    fence_mbar_init();
  }
  if (tid < 96)
    bias_sm[tid] = tid < 32 ? (a.sig_b >= 0 ? a.params[a.sig_b + tid] : 0.f)
                 : tid < 64 ? (a.gate_b >= 0 ? a.params[a.gate_b + tid - 32] : 0.f)
                            : (a.res_b >= 0 ? a.params[a.res_b + tid - 64] : 0.f);
  if (warp == 1) tmem_alloc(&tmem_base_s, 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = tmem_base_s;  // buffer ab: acc_v at ab*128 (64 cols), acc_r at ab*128 + 64 (32 cols)

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&w_full, (uint32_t)(2 * 2 * D * XB + (a.last ? 0 : R * XB)));
      tma_load_2d(wc0, &map_wc, &w_full, 0, (a.l * 2 + 0) * 2 * D);
      tma_load_2d(wc1, &map_wc, &w_full, 0, (a.l * 2 + 1) * 2 * D);
      if (!a.last) tma_load_2d(wr, &map_wr, &w_full, 0, a.l * R);
      for (int i = 0; i < n_my; ++i) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int b = tile / a.tiles_per_slot, t0 = (tile % a.tiles_per_slot) * 128;
        const int s = i % NST;
        mbar_wait(&stage_free[s], ((uint32_t)(i / NST) & 1u) ^ 1u);
        mbar_expect_tx(&in_full[s], (uint32_t)STAGE);
        tma_load_3d(smem + s * STAGE, &map_x, &in_full[s], 0, t0, b);
        tma_load_3d(smem + s * STAGE + PANEL, &map_x, &in_full[s], 0, t0 + a.dil, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(&w_full, 0);
      const uint32_t idv = make_idesc_bf16(128, 2 * D), idr = make_idesc_bf16(128, R);
      auto mma1 = [&](int i) {
        const int s = i % NST, ab = i & 1;
        const uint32_t st = smem_u32(smem + s * STAGE), av = tm + ab * 128;
        mbar_wait(&in_full[s], (uint32_t)(i / NST) & 1u);
        mbar_wait(&acc_free[ab], ((uint32_t)(i >> 1) & 1u) ^ 1u);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < R / 16; ++k)
          mma_bf16_ss(av, make_kmajor_desc(st, XB, k * 32), make_kmajor_desc(smem_u32(wc0), XB, k * 32), idv, k != 0);
#pragma unroll
        for (int k = 0; k < R / 16; ++k)
          mma_bf16_ss(av, make_kmajor_desc(st + PANEL, XB, k * 32), make_kmajor_desc(smem_u32(wc1), XB, k * 32), idv, true);
        mma_commit(&v_full[ab]);
      };
      if (n_my > 0) mma1(0);
      for (int i = 0; i < n_my; ++i) {
        if (i + 1 < n_my) mma1(i + 1);  // next tile's conv runs while this tile's gate epilogue works
        if (!a.last) {
          const int ab = i & 1;
          mbar_wait(&z_ready[ab], (uint32_t)(i >> 1) & 1u);
          tc_fence_after_sync();
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            mma_bf16_ss(tm + ab * 128 + 64, make_kmajor_desc(smem_u32(zt + ab * PANEL), XB, k * 32),
                        make_kmajor_desc(smem_u32(wr), XB, k * 32), idr, k != 0);
          mma_commit(&r_full[ab]);
        }
      }
    }
  } else {
    const int e = warp - 2, q4 = warp & 3, half = e >> 2;
    const int r = q4 * 32 + lane, c0 = 16 * half;
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const bool elected = (warp == 2 && lane == 0);
    auto ebar = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    for (int i = 0; i < n_my; ++i) {
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const int b = tile / a.tiles_per_slot, t0 = (tile % a.tiles_per_slot) * 128;
      const int s = i % NST, ab = i & 1;
      unsigned char* ztile = zt + ab * PANEL;
      unsigned char* otile = ot + ab * PANEL;
      uint32_t vs[16], vg[16], pk[8];
      mbar_wait(&v_full[ab], (uint32_t)(i >> 1) & 1u);
      tc_fence_after_sync();
      tmem_ld_32x32b_x16(tm + ab * 128 + c0 + lane_sel, vs);
      tmem_ld_32x32b_x16(tm + ab * 128 + 32 + c0 + lane_sel, vg);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float z0 = tanh_fast(__uint_as_float(vs[2 * j]) + bias_sm[c0 + 2 * j]) *
                         sigmoid_fast(__uint_as_float(vg[2 * j]) + bias_sm[32 + c0 + 2 * j]);
        const float z1 = tanh_fast(__uint_as_float(vs[2 * j + 1]) + bias_sm[c0 + 2 * j + 1]) *
                         sigmoid_fast(__uint_as_float(vg[2 * j + 1]) + bias_sm[32 + c0 + 2 * j + 1]);
        pk[j] = pack2(z0, z1);
      }
      row_store<XB, 32>(ztile, r, 32 * half, pk);
      fence_proxy_async_smem();
      tc_fence_before_sync();
      ebar();
      if (elected) {
        if (!a.last) mbar_arrive(&z_ready[ab]);
        tma_store_3d(&map_z, ztile, a.l * D, t0, b);
        tma_store_commit();
      }
      if (!a.last) {
        uint32_t vr[16], xin[8];
        mbar_wait(&in_full[s], (uint32_t)(i / NST) & 1u);  // x[t] tile (TMA) visible to this thread
        mbar_wait(&r_full[ab], (uint32_t)(i >> 1) & 1u);
        tc_fence_after_sync();
        tmem_ld_32x32b_x16(tm + ab * 128 + 64 + c0 + lane_sel, vr);
        row_load<XB, 32>(smem + s * STAGE + PANEL, r, 32 * half, xin);
        tmem_ld_wait();
        tc_fence_before_sync();
        mbar_arrive(&acc_free[ab]);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          pk[j] = pack2(__uint_as_float(vr[2 * j]) + __uint_as_float(xin[j] << 16) + bias_sm[64 + c0 + 2 * j],
                        __uint_as_float(vr[2 * j + 1]) + __uint_as_float(xin[j] & 0xffff0000u) + bias_sm[64 + c0 + 2 * j + 1]);
        row_store<XB, 32>(otile, r, 32 * half, pk);
        fence_proxy_async_smem();
        if (elected) tma_store_wait_read<1>();  // every store but this tile's z store has finished reading smem
        ebar();
        if (elected) {
          tma_store_3d(&map_xout, otile, 0, a.dil_next + t0, b);
          tma_store_commit();
          mbar_arrive(&stage_free[s]);
        }
      } else {
        tc_fence_before_sync();
        mbar_arrive(&acc_free[ab]);
        if (elected) tma_store_wait_read<1>();
        ebar();
        if (elected) mbar_arrive(&stage_free[s]);
      }
    }
    if (elected) tma_store_wait_all<0>();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tm, 256);
}

struct LayerDxPArgs {
  int dil, l, has_next, n_tiles, tiles_per_slot;
};

template <int R, int D>
__global__ void __launch_bounds__(320, 1)
k_layer_bwd_dx_p_umma(const __grid_constant__ CUtensorMap map_dv, const __grid_constant__ CUtensorMap map_wd,
                      const __grid_constant__ CUtensorMap map_dxn, const __grid_constant__ CUtensorMap map_dxo,
                      LayerDxPArgs a) {
  static_assert(R == 32 && D == 32, "64-byte activation rows");
  constexpr int XB = 64, VB = 128;
  constexpr int PANEL = 128 * XB, VT = 128 * VB;
  constexpr int STAGE = 2 * VT + PANEL;  // dv[t] | dv[t+dil] | dx'
  constexpr int NST = 4;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  unsigned char* ot = smem + NST * STAGE;  // [3] output tiles (a store may still be reading two tiles back)
  unsigned char* wd0 = ot + 3 * PANEL;
  unsigned char* wd1 = wd0 + R * VB;
  __shared__ __align__(8) uint64_t w_full, in_full[NST], stage_free[NST], a_full[2], acc_free[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_my = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  if (tid == 0) {
    mbar_init(&w_full, 1);
    for (int i = 0; i < NST; ++i) {
      mbar_init(&in_full[i], 1);
      mbar_init(&stage_free[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&acc_free[i], 256);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, 64);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tm = tmem_base_s;  // buffer ab at ab*32
  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(&w_full, (uint32_t)(2 * R * VB));
      tma_load_2d(wd0, &map_wd, &w_full, 0, (a.l * 2 + 0) * R);
      tma_load_2d(wd1, &map_wd, &w_full, 0, (a.l * 2 + 1) * R);
      for (int i = 0; i < n_my; ++i) {
        const int tile = (int)blockIdx.x + i * (int)gridDim.x;
        const int b = tile / a.tiles_per_slot, t0 = (tile % a.tiles_per_slot) * 128;
        const int s = i % NST;
        unsigned char* st = smem + s * STAGE;
        mbar_wait(&stage_free[s], ((uint32_t)(i / NST) & 1u) ^ 1u);
        mbar_expect_tx(&in_full[s], (uint32_t)(2 * VT + (a.has_next ? PANEL : 0)));
        tma_load_3d(st, &map_dv, &in_full[s], 0, t0, b);
        tma_load_3d(st + VT, &map_dv, &in_full[s], 0, t0 + a.dil, b);  // rows >= T: zero fill == truncated gradient
        if (a.has_next) tma_load_3d(st + 2 * VT, &map_dxn, &in_full[s], 0, t0, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(&w_full, 0);
      const uint32_t idesc = make_idesc_bf16(128, R);
      for (int i = 0; i < n_my; ++i) {
        const int s = i % NST, ab = i & 1;
        const uint32_t st = smem_u32(smem + s * STAGE);
        mbar_wait(&in_full[s], (uint32_t)(i / NST) & 1u);
        mbar_wait(&acc_free[ab], ((uint32_t)(i >> 1) & 1u) ^ 1u);
        tc_fence_after_sync();
#pragma unroll
        for (int k = 0; k < 2 * D / 16; ++k)  // dv[t] . W[1]^T
          mma_bf16_ss(tm + ab * 32, make_kmajor_desc(st, VB, k * 32), make_kmajor_desc(smem_u32(wd1), VB, k * 32), idesc, k != 0);
#pragma unroll
        for (int k = 0; k < 2 * D / 16; ++k)  // dv[t+dil] . W[0]^T
          mma_bf16_ss(tm + ab * 32, make_kmajor_desc(st + VT, VB, k * 32), make_kmajor_desc(smem_u32(wd0), VB, k * 32), idesc, true);
        mma_commit(&a_full[ab]);
      }
    }
  } else {
    const int e = warp - 2, q4 = warp & 3, half = e >> 2;
    const int r = q4 * 32 + lane, c0 = 16 * half;
    const uint32_t lane_sel = (uint32_t)(q4 * 32) << 16;
    const bool elected = (warp == 2 && lane == 0);
    auto ebar = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    for (int i = 0; i < n_my; ++i) {
      const int tile = (int)blockIdx.x + i * (int)gridDim.x;
      const int b = tile / a.tiles_per_slot, t0 = (tile % a.tiles_per_slot) * 128;
      const int s = i % NST, ab = i & 1;
      unsigned char* otile = ot + (i % 3) * PANEL;
      uint32_t vr[16], xin[8], pk[8];
      mbar_wait(&in_full[s], (uint32_t)(i / NST) & 1u);
      mbar_wait(&a_full[ab], (uint32_t)(i >> 1) & 1u);
      tc_fence_after_sync();
      tmem_ld_32x32b_x16(tm + ab * 32 + c0 + lane_sel, vr);
      if (a.has_next) {
        row_load<XB, 32>(smem + s * STAGE + 2 * VT, r, 32 * half, xin);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) xin[j] = 0u;
      }
      tmem_ld_wait();
      tc_fence_before_sync();
      mbar_arrive(&acc_free[ab]);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        pk[j] = pack2(__uint_as_float(vr[2 * j]) + __uint_as_float(xin[j] << 16),
                      __uint_as_float(vr[2 * j + 1]) + __uint_as_float(xin[j] & 0xffff0000u));
      row_store<XB, 32>(otile, r, 32 * half, pk);
      fence_proxy_async_smem();
      if (elected) tma_store_wait_read<1>();  // all stores but the previous tile's have drained: buffer (i+1)%3 is free
      ebar();
      if (elected) {
        tma_store_3d(&map_dxo, otile, 0, t0, b);
        tma_store_commit();
        mbar_arrive(&stage_free[s]);
      }
    }
    if (elected) tma_store_wait_all<0>();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tm, 64);
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

struct LayerMaps {
  const void* ws = nullptr;
  const void* model = nullptr;
  int T = -1;
  std::vector<CUtensorMap> x;  // per layer: xfull_l [B][dil+T][R]
  CUtensorMap z, wc, wr, wd, dv, dx[2], dz, wrn;
};

bool umma_layer_supported(const wn_model* m) {
  static const bool disabled = getenv("WN_DISABLE_UMMA") != nullptr || getenv("WN_DISABLE_UMMA_LAYER") != nullptr;
  const wn_arch& a = m->a;
  return !disabled && a.n_res == 32 && a.n_dil == 32;
}

static LayerMaps* get_maps(wn_model* m, unsigned char* ws, int T, int* rc) {
  static thread_local LayerMaps cache;  // one model per process in practice; re-encoded when ws/T change
  *rc = WN_OK;
  if (cache.model == m && cache.ws == ws && cache.T == T && (int)cache.x.size() == m->L) return &cache;
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
  if ((*rc = map2ds(&cache.wd, ws + wl.wdT, 2 * D, (uint64_t)m->L * 2 * R, (uint32_t)(2 * D), (uint32_t)R, (int)D * 4))) return nullptr;
  if ((*rc = map3d(&cache.dv, ws + wl.dv, 2 * D, (uint64_t)T, B, (uint32_t)(2 * D), 128, (int)D * 4))) return nullptr;
  for (int i = 0; i < 2; ++i)
    if ((*rc = map3d(&cache.dx[i], ws + wl.dx[i], R, (uint64_t)T, B, (uint32_t)R, 128, (int)R * 2))) return nullptr;
  if ((*rc = map3d(&cache.dz, ws + wl.dz, LD, (uint64_t)T, B, (uint32_t)D, 128, (int)D * 2))) return nullptr;
  if ((*rc = map2ds(&cache.wrn, ws + wl.wrN, R, (uint64_t)m->L * D, (uint32_t)R, (uint32_t)D, (int)R * 2))) return nullptr;
  cache.ws = ws;
  cache.T = T;
  cache.model = m;
  return &cache;
}

int launch_prep_layer_umma(wn_model* m, const float* d_params, unsigned char* ws, cudaStream_t st) {
  const WorkspaceLayout& wl = m->wl;
  k_prep_layer_weights<<<m->L, 256, 0, st>>>(d_params, m->d_layers, m->L, m->a.n_res, m->a.n_dil,
                                             reinterpret_cast<bf16*>(ws + wl.wcT), reinterpret_cast<bf16*>(ws + wl.wrT),
                                             reinterpret_cast<bf16*>(ws + wl.wdT), reinterpret_cast<bf16*>(ws + wl.wrN));
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
  const dim3 grid((T + 127) / 128, m->n_slots);
  if (a.n_gc_embed > 0) {  // global conditioning: per-tile variant (reads the per-id projection table)
    LayerFwdUmmaArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.params = d_params;
    fa.sig_b = ld.sig_b; fa.gate_b = ld.gate_b; fa.res_b = ld.res_b;
    fa.gc_tbl = reinterpret_cast<const float*>(ws + wl.gc_tbl) + (size_t)l * C1 * 2 * a.n_dil;
    fa.ids = d_ids;
    fa.T = T; fa.dil = ld.dil; fa.l = l; fa.C1 = C1;
    fa.last = last;
    fa.dil_next = last ? 0 : m->layers[l + 1].dil;
    const size_t smem = 3 * 128 * 64 + 2 * 64 * 64 + 32 * 64 + 1024;
    ProfScope ps(PROF_LAYER_FWD, st);
    k_layer_fwd_umma<32, 32><<<grid, 128, smem, st>>>(mp->x[l], mxo, mp->z, mp->wc, mp->wr, fa);
    WN_LAUNCH_CHECK();
    return WN_OK;
  }
  LayerFwdPArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.params = d_params;
  pa.sig_b = ld.sig_b; pa.gate_b = ld.gate_b; pa.res_b = ld.res_b;
  pa.T = T; pa.dil = ld.dil; pa.l = l; pa.last = last;
  pa.dil_next = last ? 0 : m->layers[l + 1].dil;
  pa.tiles_per_slot = (T + 127) / 128;
  pa.n_tiles = pa.tiles_per_slot * m->n_slots;
  const size_t smem = 4 * 2 * 8192 + 4 * 8192 + 2 * 64 * 64 + 32 * 64 + 1024;
  WN_CUDA_CHECK(cudaFuncSetAttribute(k_layer_fwd_p_umma<32, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nblk = persist_grid(std::max(1, std::min(pa.n_tiles, 2 * m->sm_count)));
  ProfScope ps(PROF_LAYER_FWD, st);
  k_layer_fwd_p_umma<32, 32><<<nblk, 320, smem, st>>>(mp->x[l], mxo, mp->z, mp->wc, mp->wr, pa);
  WN_LAUNCH_CHECK();
  return WN_OK;
}

// dx_out = dxbuf[l & 1], dx_next = dxbuf[(l + 1) & 1] (as in the generation-1 orchestration)
int launch_layer_bwd_dx_umma(wn_model* m, unsigned char* ws, int T, int l, cudaStream_t st) {
  int rc;
  LayerMaps* mp = get_maps(m, ws, T, &rc);
  if (!mp) return rc;
  LayerDxPArgs da;
  da.dil = m->layers[l].dil;
  da.l = l;
  da.has_next = (l + 1 < m->L);
  da.tiles_per_slot = (T + 127) / 128;
  da.n_tiles = da.tiles_per_slot * m->n_slots;
  const size_t smem = 4 * (2 * 16384 + 8192) + 3 * 8192 + 2 * 32 * 128 + 1024;
  WN_CUDA_CHECK(cudaFuncSetAttribute(k_layer_bwd_dx_p_umma<32, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nblk = persist_grid(std::max(1, std::min(da.n_tiles, m->sm_count)));
  ProfScope ps(PROF_LAYER_BWD_B, st);
  k_layer_bwd_dx_p_umma<32, 32><<<nblk, 320, smem, st>>>(mp->dv, mp->wd, mp->dx[(l + 1) & 1], mp->dx[l & 1], da);
  WN_LAUNCH_CHECK();
  return WN_OK;
}

}  // namespace wn

namespace wn {

bool umma_gate_supported(const wn_model* m) { return umma_layer_supported(m) && m->a.n_gc_embed == 0; }

// gate backward + conv / residual weight and bias gradients of layer l (dx_next = dxbuf[(l+1)&1])
int launch_layer_bwd_gate_umma(wn_model* m, const float* d_params, unsigned char* ws, int T, int l, float* d_grads,
                               cudaStream_t st) {
  int rc;
  LayerMaps* mp = get_maps(m, ws, T, &rc);
  if (!mp) return rc;
  const LayerDesc& ld = m->layers[l];
  LayerGateUmmaArgs ga;
  memset(&ga, 0, sizeof(ga));
  ga.params = d_params;
  ga.grads = d_grads;
  ga.sig = ld.sig; ga.gate = ld.gate; ga.res = ld.res;
  ga.sig_b = ld.sig_b; ga.gate_b = ld.gate_b; ga.res_b = ld.res_b;
  ga.T = T; ga.dil = ld.dil; ga.l = l;
  ga.has_next = (l + 1 < m->L);
  ga.tiles_per_slot = (T + 127) / 128;
  ga.n_tiles = ga.tiles_per_slot * m->n_slots;
  const size_t smem = 3 * (6 * 128 * 64 + 128 * 128) + 2 * 64 * 64 + 32 * 64 + 1024;
  WN_CUDA_CHECK(cudaFuncSetAttribute(k_layer_bwd_gate_umma<32, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = persist_grid(std::max(1, std::min(ga.n_tiles, m->sm_count)));
  ProfScope ps(PROF_LAYER_BWD_A, st);
  k_layer_bwd_gate_umma<32, 32><<<grid, 320, smem, st>>>(mp->x[l], mp->dz, mp->dx[(l + 1) & 1], mp->dv, mp->wc, mp->wrn, ga);
  WN_LAUNCH_CHECK();
  return WN_OK;
}

}  // namespace wn
