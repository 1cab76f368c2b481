// Internal model description shared by the host-side registry and the kernel launchers.
// Mirrors the reference's variable registry (arch.py:85-103,112-142) as a flat-arena layout.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/wavenet_b200.h"

namespace wn {

constexpr int WN_EMBED_PARTS = 512;

struct ParamEntry {
  std::string name;
  int64_t offset;  // element offset into the fp32 arena
  int32_t ndim;
  int64_t shape[3];
  int32_t kind;  // WN_KIND_*
  int64_t numel() const {
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i) n *= shape[i];
    return n;
  }
};

// Per-layer offsets (elements, fp32 arena) -- plain-old-data, passed to kernels by value
// and also uploaded as a device table.  -1 == tensor absent (no bias / no GC).
struct LayerDesc {
  int64_t sig, sig_b, gate, gate_b, gc_sig, gc_gate, res, res_b, skip, skip_b;
  int64_t lc_sig, lc_gate;  // LC_SIGNAL / LC_GATE [n_lc_out][D] (reference tmodel.py:156-160), -1 without local conditioning
  int64_t save_off;   // element offset into the bf16 SAVE arena, layout [n_slots][dil][R]
  int64_t xfull_off;  // byte offset of xfull_l in the workspace (filled per slice_sz)
  int32_t dil;
  int32_t pad;
};

// Workspace carve-up for a given slice_sz (byte offsets)
struct WorkspaceLayout {
  int32_t T = -1;
  int64_t wbf;        // bf16 mirror of the parameter arena
  int64_t z;          // [B*T][L*D] bf16
  int64_t h1, h2;     // [B*T][S], [B*T][P] bf16
  int64_t hm1, hm2;   // relu masks of h1 / h2 as bits: [ceil(B*T / 128) * 128][8] uint32 (bit c % 32 of word c / 32 = h[.][c] > 0),
                      // written by the fused post-net forward, read by its backward instead of the activations themselves
  int64_t dlogits;    // [B*T][Q] bf16
  int64_t dp1, dskip; // [B*T][P], [B*T][S] bf16
  int64_t dz;         // [L][B*T][D] bf16: per-layer planes (dense 64-byte rows for the layer backward)
  int64_t dv;         // [B*T][2D] bf16
  int64_t dx[2];      // [B*T][R] bf16 ping/pong: data gradient (GC path) / its Y part (fused backward)
  int64_t gc_tbl;     // [L][C+1][2D] fp32 (GC projections incl. nothing else)
  int64_t dgc_tbl;    // same shape, gradient
  int64_t skip_bias;  // [S] fp32, sum over layers of SKIP_BIAS
  int64_t tile_ctr;   // [L][4] int32: dynamic tile-scheduler counters of the persistent layer kernels (fwd: 0,1; bwd: 2,3)
  int64_t embed_part; // [WN_EMBED_PARTS][Q + 1][R] fp32: per-CTA partial PRE gradients
  // transposed / concatenated bf16 weight copies for the tcgen05 kernels (B operands, N x K K-major)
  int64_t wsT;        // [S][L*D]   = SKIP_l[d][s] at [s][l*D+d]
  int64_t wsCat;      // [L*D][S]   = SKIP_l stacked over layers
  int64_t w1T;        // [P][S]     = POST1^T
  int64_t w2T;        // [Q][P]     = POST2^T
  int64_t wcT, wrT, wrN;     // per-layer conv / residual operands of the tcgen05 layer kernels (layer_umma.cu)
  int64_t wdP;               // [L][R][4D]: B operand of the wide layers' data gradient (train_umma.cu)
  // ---- local conditioning (train_umma.cu "local conditioning"): every activation row is 128 bf16 wide (LCP), channels
  // beyond n_lc_in / n_lc_out are zero.  lc_x[0] = mel frames, lc_x[i + 1] = output of upsampling level i.
  int64_t lc_x[9];           // [B * T_i][128] bf16, T_i = T / hop * prod(s_0 .. s_{i-1})
  int64_t lc_dx[9];          // gradients of the same (index 0 unused)
  int64_t cond;              // [L][B*T][2D] bf16: lc_up . [LC_SIGNAL_l | LC_GATE_l]
  int64_t dcond;             // [L][B*T][2D] bf16: its gradient, dv_l, written by the layer backward
  int64_t lc_wup[8];         // [s_i * 128][128] bf16: B operand of level i  (row k * 128 + o, column c) = LC_UPSAMPLE_i[k][o][c]
  int64_t lc_wupT[8];        // [128][s_i * 128] bf16: its transpose, B operand of the level's data gradient
  int64_t lc_wcat;           // [L * 2D][128] bf16: row l * 2D + n = (LC_SIGNAL_l | LC_GATE_l)[c][n]
  int64_t lc_wcatT;          // [128][L * 2D] bf16
  int64_t lc_gtmp;           // fp32 scratch for the weight gradients: [L][128][2D] | per level [128][s_i * 128]
  int64_t total;
  std::vector<int64_t> xfull;  // per layer: [B][dil+T][R] bf16
};

}  // namespace wn

struct wn_model {
  wn_arch a;
  uint64_t serial;  // unique per wn_model_create in this process: keys caches that must not survive a reused address
  int32_t n_slots;
  int32_t L;
  std::vector<wn::ParamEntry> params;
  int64_t n_param_elems;
  int64_t off_pre, off_pre_b, off_gc_embed, off_post1, off_post1_b, off_post2, off_post2_b;
  int64_t off_lc_up[8];  // LC_UPSAMPLE_i [s_i][n_lc_out][n_lc_in | n_lc_out]
  int32_t lc_hop;        // prod(lc_upsample), 1 without local conditioning
  std::vector<wn::LayerDesc> layers;
  int64_t save_elems;
  wn::WorkspaceLayout wl;  // cached for the last slice_sz
  // lazily created device-side tables (owned by the handle)
  wn::LayerDesc* d_layers = nullptr;
  int32_t d_layers_T = -1;
  uint8_t* d_kind = nullptr;  // per arena element: 1 == L2-regularised filter
  int sm_count = 0;
};

namespace wn {
void set_error(const char* fmt, ...);
const WorkspaceLayout& workspace_layout(wn_model* m, int32_t T);
inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }
}  // namespace wn
