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
  unsigned char* otile = ztile + Z_TILE;
  unsigned char* wc0 = otile + X_TILE;
  unsigned char* wc1 = wc0 + WC_TILE;
  unsigned char* wr = wc1 + WC_TILE;
  __shared__ __align__(8) uint64_t bar_in, bar_v, bar_r;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = blockIdx.y, t0 = blockIdx.x * 128;
  constexpr uint32_t NCOL = (2 * D + R) <= 128 ? 128 : 256;

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
  const uint32_t acc_v = tmem_base_s, acc_r = tmem_base_s + 2 * D;

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
  unsigned char* otile = dxn + X_TILE;
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

struct LayerMaps {
  const void* ws = nullptr;
  const void* model = nullptr;
  int T = -1;
  std::vector<CUtensorMap> x;  // per layer: xfull_l [B][dil+T][R]
  CUtensorMap z, wc, wr, wd, dv, dx[2];
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
  LayerFwdUmmaArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.params = d_params;
  fa.sig_b = ld.sig_b; fa.gate_b = ld.gate_b; fa.res_b = ld.res_b;
  const int C1 = a.n_gc_category + 1;
  fa.gc_tbl = a.n_gc_embed > 0 ? reinterpret_cast<const float*>(ws + wl.gc_tbl) + (size_t)l * C1 * 2 * a.n_dil : nullptr;
  fa.ids = d_ids;
  fa.T = T; fa.dil = ld.dil; fa.l = l; fa.C1 = C1;
  fa.last = (l + 1 == m->L);
  fa.dil_next = fa.last ? 0 : m->layers[l + 1].dil;
  const CUtensorMap& mxo = fa.last ? mp->x[l] : mp->x[l + 1];
  const size_t smem = 4 * 128 * 64 + 2 * 64 * 64 + 32 * 64 + 1024;
  const dim3 grid((T + 127) / 128, m->n_slots);
  ProfScope ps(PROF_LAYER_FWD, st);
  k_layer_fwd_umma<32, 32><<<grid, 128, smem, st>>>(mp->x[l], mxo, mp->z, mp->wc, mp->wr, fa);
  WN_LAUNCH_CHECK();
  return WN_OK;
}

// dx_out = dxbuf[l & 1], dx_next = dxbuf[(l + 1) & 1] (as in the generation-1 orchestration)
int launch_layer_bwd_dx_umma(wn_model* m, unsigned char* ws, int T, int l, cudaStream_t st) {
  int rc;
  LayerMaps* mp = get_maps(m, ws, T, &rc);
  if (!mp) return rc;
  LayerDxUmmaArgs da;
  da.dil = m->layers[l].dil;
  da.l = l;
  da.has_next = (l + 1 < m->L);
  const size_t smem = 2 * 128 * 128 + 2 * 32 * 128 + 2 * 128 * 64 + 1024;
  const dim3 grid((T + 127) / 128, m->n_slots);
  WN_CUDA_CHECK(cudaFuncSetAttribute(k_layer_bwd_dx_umma<32, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope ps(PROF_LAYER_BWD_B, st);
  k_layer_bwd_dx_umma<32, 32><<<grid, 128, smem, st>>>(mp->dv, mp->wd, mp->dx[(l + 1) & 1], mp->dx[l & 1], da);
  WN_LAUNCH_CHECK();
  return WN_OK;
}

}  // namespace wn
