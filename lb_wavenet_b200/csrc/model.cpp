// Host-only part of the C ABI: model handle, variable registry, workspace carve-up.
// Needs no GPU (CPU tests call it to check names/shapes against the reference contract).
#include "model.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace wn {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static const int64_t kArenaAlign = 64;  // elements (256 B for fp32, 128 B for the bf16 mirror)

static int64_t add_param(wn_model* m, const std::string& name, std::initializer_list<int64_t> shape,
                         int32_t kind) {
  ParamEntry e;
  e.name = name;
  e.offset = m->n_param_elems;
  e.ndim = (int32_t)shape.size();
  e.shape[0] = e.shape[1] = e.shape[2] = 1;
  int i = 0;
  for (int64_t s : shape) e.shape[i++] = s;
  e.kind = kind;
  m->params.push_back(e);
  m->n_param_elems = align_up(m->n_param_elems + e.numel(), kArenaAlign);
  return e.offset;
}

const WorkspaceLayout& workspace_layout(wn_model* m, int32_t T) {
  if (m->wl.T == T) return m->wl;
  WorkspaceLayout w;
  const wn_arch& a = m->a;
  const int64_t B = m->n_slots, L = m->L, R = a.n_res, D = a.n_dil, S = a.n_skip, P = a.n_post,
                Q = a.n_quant;
  const int64_t rows = B * (int64_t)T;
  int64_t off = 0;
  auto take = [&](int64_t bytes) {
    int64_t o = off;
    off = align_up(off + bytes, 1024);
    return o;
  };
  w.T = T;
  w.wbf = take(m->n_param_elems * 2);
  w.xfull.resize(L);
  for (int64_t l = 0; l < L; ++l) w.xfull[l] = take(B * (m->layers[l].dil + (int64_t)T) * R * 2);
  w.z = take(rows * L * D * 2);
  w.h1 = take(rows * S * 2);
  w.h2 = take(rows * P * 2);
  w.hm1 = take(align_up(rows, 128) * 32);
  w.hm2 = take(align_up(rows, 128) * 32);
  w.dlogits = take(rows * Q * 2);
  w.dp1 = take(rows * P * 2);
  w.dskip = take(rows * S * 2);
  w.dz = take(rows * L * D * 2);
  w.dv = take(rows * 2 * D * 2);
  w.dx[0] = take(rows * R * 2);
  w.dx[1] = take(rows * R * 2);
  const int64_t gc_elems = a.n_gc_embed > 0 ? L * (int64_t)(a.n_gc_category + 1) * 2 * D : 0;
  w.gc_tbl = take(gc_elems * 4);
  w.dgc_tbl = take(gc_elems * 4);
  w.skip_bias = take(S * 4);
  w.tile_ctr = take(L * 4 * 4);
  w.embed_part = take((int64_t)WN_EMBED_PARTS * (Q + 1) * R * 4);
  w.wsT = take(S * L * D * 2);
  w.wsCat = take(L * D * S * 2);
  w.w1T = take(P * S * 2);
  w.w2T = take(Q * P * 2);
  w.wcT = take(L * 2 * 2 * D * R * 2);
  w.wrT = take(L * R * D * 2);
  w.wrN = take(L * D * R * 2);
  w.wdP = take(L * R * 4 * D * 2);
  for (int i = 0; i < 9; ++i) w.lc_x[i] = w.lc_dx[i] = 0;
  for (int i = 0; i < 8; ++i) w.lc_wup[i] = w.lc_wupT[i] = 0;
  w.cond = w.dcond = w.lc_wcat = w.lc_wcatT = w.lc_gtmp = 0;
  if (a.n_lc_out > 0) {
    const int64_t LCP = 128;
    int64_t r = rows / m->lc_hop, gt = L * LCP * 2 * D;
    for (int i = 0; i <= a.n_lc_layers; ++i) {
      w.lc_x[i] = take(align_up(r, 128) * LCP * 2);
      if (i > 0) w.lc_dx[i] = take(align_up(r, 128) * LCP * 2);
      if (i < a.n_lc_layers) {
        w.lc_wup[i] = take((int64_t)a.lc_upsample[i] * LCP * LCP * 2);
        w.lc_wupT[i] = take((int64_t)a.lc_upsample[i] * LCP * LCP * 2);
        gt += LCP * a.lc_upsample[i] * LCP;
        r *= a.lc_upsample[i];
      }
    }
    w.cond = take(L * rows * 2 * D * 2);
    w.dcond = take(L * rows * 2 * D * 2);
    w.lc_wcat = take(L * 2 * D * LCP * 2);
    w.lc_wcatT = take(L * 2 * D * LCP * 2);
    w.lc_gtmp = take(gt * 4);
  }
  w.total = off;
  m->wl = w;
  return m->wl;
}

}  // namespace wn

using namespace wn;

extern "C" {

int32_t wn_abi_version(void) { return WN_ABI_VERSION; }

// CRC-32C (Castagnoli, reflected polynomial 0x82F63B78), slicing-by-8: the checksum of TensorFlow's tensor-bundle
// checkpoint files (lb_wavenet_b200/tfbundle.py).  Host memory; `crc` is the running value (0 to start).
uint32_t wn_crc32c(uint32_t crc, const void* h_data, uint64_t n) {
  static uint32_t tbl[8][256];
  static bool ready = false;
  if (!ready) {
    for (uint32_t i = 0; i < 256; ++i) {
      uint32_t c = i;
      for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      tbl[0][i] = c;
    }
    for (uint32_t i = 0; i < 256; ++i)
      for (int t = 1; t < 8; ++t) tbl[t][i] = (tbl[t - 1][i] >> 8) ^ tbl[0][tbl[t - 1][i] & 0xffu];
    ready = true;
  }
  const unsigned char* p = static_cast<const unsigned char*>(h_data);
  uint32_t c = ~crc;
  while (n >= 8) {
    uint32_t lo, hi;
    memcpy(&lo, p, 4);
    memcpy(&hi, p + 4, 4);
    lo ^= c;
    c = tbl[7][lo & 0xffu] ^ tbl[6][(lo >> 8) & 0xffu] ^ tbl[5][(lo >> 16) & 0xffu] ^ tbl[4][lo >> 24] ^
        tbl[3][hi & 0xffu] ^ tbl[2][(hi >> 8) & 0xffu] ^ tbl[1][(hi >> 16) & 0xffu] ^ tbl[0][hi >> 24];
    p += 8;
    n -= 8;
  }
  while (n--) c = (c >> 8) ^ tbl[0][(c ^ *p++) & 0xffu];
  return ~c;
}

const char* wn_last_error(void) { return g_err; }

int wn_model_create(const wn_arch* arch, int32_t n_slots, wn_model** out) {
  if (!arch || !out) {
    set_error("wn_model_create: null argument");
    return WN_ERR_INVALID;
  }
  const wn_arch& a = *arch;
  auto bad = [&](const char* what) {
    set_error("wn_model_create: unsupported architecture: %s", what);
    return (int)WN_ERR_UNSUPPORTED;
  };
  if (a.n_quant != 256) return bad("n_quant must be 256");
  if (a.n_blocks < 1 || a.n_block_layers < 1 || a.n_block_layers > 14) return bad("n_blocks/n_block_layers");
  if (a.n_res % 16 || a.n_res < 16 || a.n_res > 128) return bad("n_res must be a multiple of 16 in [16,128]");
  if (a.n_dil % 16 || a.n_dil < 16 || a.n_dil > 128) return bad("n_dil must be a multiple of 16 in [16,128]");
  if (a.n_skip % 16 || a.n_skip < 16 || a.n_skip > 512) return bad("n_skip must be a multiple of 16 in [16,512]");
  if (a.n_post % 16 || a.n_post < 16 || a.n_post > 512) return bad("n_post must be a multiple of 16 in [16,512]");
  if (a.n_gc_embed < 0 || a.n_gc_embed > 64) return bad("n_gc_embed must be in [0,64]");
  if (a.n_gc_embed > 0 && a.n_gc_category < 1) return bad("n_gc_category must be >= 1 with global conditioning");
  if (n_slots < 1) return bad("n_slots must be >= 1");
  if (a.n_lc_out < 0 || a.n_lc_out > 128 || a.n_lc_in < 0 || a.n_lc_in > 128) return bad("n_lc_in / n_lc_out must be in [0,128]");
  int64_t hop = 1;
  if (a.n_lc_out > 0) {
    if (a.n_lc_in < 1) return bad("local conditioning needs n_lc_in >= 1");
    if (a.n_lc_layers < 1 || a.n_lc_layers > 8) return bad("lc_upsample must hold 1..8 strides");
    for (int i = 0; i < a.n_lc_layers; ++i) {
      if (a.lc_upsample[i] < 1 || a.lc_upsample[i] > 16) return bad("lc_upsample strides must be in [1,16]");
      hop *= a.lc_upsample[i];
    }
    if (hop > 65536) return bad("prod(lc_upsample) too large");
    if (a.n_res != 32 || a.n_dil != 32) return bad("local conditioning is built for n_res == n_dil == 32 (the fused layer kernels)");
  }

  wn_model* m = new wn_model();
  {
    static std::atomic<uint64_t> next_serial{1};
    m->serial = next_serial.fetch_add(1);
  }
  m->a = a;
  m->n_slots = n_slots;
  m->L = a.n_blocks * a.n_block_layers;
  m->n_param_elems = 0;
  m->save_elems = 0;
  const int64_t R = a.n_res, D = a.n_dil, S = a.n_skip, P = a.n_post, Q = a.n_quant, G = a.n_gc_embed;
  const bool gc = G > 0, ub = a.use_bias != 0;
  // registration order == graph construction order of the reference (tmodel.py:292-328)
  m->off_gc_embed = gc ? add_param(m, "GC_EMBED", {a.n_gc_category + 1, G}, WN_KIND_FILTER) : -1;
  m->off_pre = add_param(m, "PRE", {Q, R}, WN_KIND_FILTER);
  m->off_pre_b = ub ? add_param(m, "PRE_BIAS", {R}, WN_KIND_BIAS) : -1;
  const bool lc = a.n_lc_out > 0;
  m->lc_hop = (int32_t)hop;
  for (int i = 0; i < 8; ++i) m->off_lc_up[i] = -1;
  if (lc)  // tmodel.py:68-83 (_preprocess_lc runs right after _preprocess, tmodel.py:307-311); shape arch.py:75-80
    for (int i = 0; i < a.n_lc_layers; ++i) {
      char nm[32];
      snprintf(nm, sizeof(nm), "LC_UPSAMPLE_%d", i);
      m->off_lc_up[i] = add_param(m, nm, {(int64_t)a.lc_upsample[i], (int64_t)a.n_lc_out, (int64_t)(i == 0 ? a.n_lc_in : a.n_lc_out)}, WN_KIND_FILTER);
    }
  for (int b = 0; b < a.n_blocks; ++b) {
    for (int bl = 0; bl < a.n_block_layers; ++bl) {
      char sfx[32];
      snprintf(sfx, sizeof(sfx), "_%d_%d", b, bl);
      LayerDesc d;
      memset(&d, 0, sizeof(d));
      d.dil = 1 << bl;  // tmodel.py:318
      d.save_off = m->save_elems;
      m->save_elems += (int64_t)n_slots * d.dil * R;
      d.sig = add_param(m, std::string("SIGNAL") + sfx, {2, R, D}, WN_KIND_FILTER);
      d.sig_b = ub ? add_param(m, std::string("SIGNAL_BIAS") + sfx, {D}, WN_KIND_BIAS) : -1;
      d.gate = add_param(m, std::string("GATE") + sfx, {2, R, D}, WN_KIND_FILTER);
      d.gate_b = ub ? add_param(m, std::string("GATE_BIAS") + sfx, {D}, WN_KIND_BIAS) : -1;
      d.gc_sig = gc ? add_param(m, std::string("GC_SIGNAL") + sfx, {G, D}, WN_KIND_FILTER) : -1;
      d.gc_gate = gc ? add_param(m, std::string("GC_GATE") + sfx, {G, D}, WN_KIND_FILTER) : -1;
      d.lc_sig = lc ? add_param(m, std::string("LC_SIGNAL") + sfx, {(int64_t)a.n_lc_out, D}, WN_KIND_FILTER) : -1;  // tmodel.py:156-160
      d.lc_gate = lc ? add_param(m, std::string("LC_GATE") + sfx, {(int64_t)a.n_lc_out, D}, WN_KIND_FILTER) : -1;
      d.res = add_param(m, std::string("RESIDUAL") + sfx, {D, R}, WN_KIND_FILTER);
      d.res_b = ub ? add_param(m, std::string("RESIDUAL_BIAS") + sfx, {R}, WN_KIND_BIAS) : -1;
      d.skip = add_param(m, std::string("SKIP") + sfx, {D, S}, WN_KIND_FILTER);
      d.skip_b = ub ? add_param(m, std::string("SKIP_BIAS") + sfx, {S}, WN_KIND_BIAS) : -1;
      m->layers.push_back(d);
    }
  }
  m->off_post1 = add_param(m, "POST1", {S, P}, WN_KIND_FILTER);
  m->off_post1_b = ub ? add_param(m, "POST1_BIAS", {P}, WN_KIND_BIAS) : -1;
  m->off_post2 = add_param(m, "POST2", {P, Q}, WN_KIND_FILTER);
  m->off_post2_b = ub ? add_param(m, "POST2_BIAS", {Q}, WN_KIND_BIAS) : -1;
  *out = m;
  return WN_OK;
}

int32_t wn_n_layers(const wn_model* m) { return m->L; }

int32_t wn_recep_field(const wn_model* m) {
  // reference tmodel.py:50-51
  return m->a.n_blocks * ((1 << m->a.n_block_layers) - 1);
}

int32_t wn_param_count(const wn_model* m) { return (int32_t)m->params.size(); }

int64_t wn_param_elems(const wn_model* m) { return m->n_param_elems; }

int wn_param_info(const wn_model* m, int32_t i, char* name, int32_t name_cap, int64_t* offset,
                  int32_t* ndim, int64_t* shape3, int32_t* kind) {
  if (i < 0 || i >= (int32_t)m->params.size()) {
    set_error("wn_param_info: index %d out of range", i);
    return WN_ERR_INVALID;
  }
  const ParamEntry& e = m->params[i];
  if (name && name_cap > 0) {
    strncpy(name, e.name.c_str(), name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (offset) *offset = e.offset;
  if (ndim) *ndim = e.ndim;
  if (shape3) {
    shape3[0] = e.shape[0];
    shape3[1] = e.shape[1];
    shape3[2] = e.shape[2];
  }
  if (kind) *kind = e.kind;
  return WN_OK;
}

int64_t wn_save_elems(const wn_model* m) { return m->save_elems; }

int wn_save_info(const wn_model* m, int32_t layer, int64_t* offset, int32_t* dil) {
  if (layer < 0 || layer >= m->L) {
    set_error("wn_save_info: layer %d out of range", layer);
    return WN_ERR_INVALID;
  }
  if (offset) *offset = m->layers[layer].save_off;
  if (dil) *dil = m->layers[layer].dil;
  return WN_OK;
}

int64_t wn_workspace_bytes(const wn_model* m, int32_t slice_sz) {
  if (slice_sz < 2) {
    set_error("wn_workspace_bytes: slice_sz must be >= 2");
    return WN_ERR_INVALID;
  }
  if (m->a.n_lc_out > 0 && slice_sz % m->lc_hop != 0) {
    set_error("wn_workspace_bytes: slice_sz %d must be a multiple of prod(lc_upsample) = %d (data.py:32-37)", slice_sz, m->lc_hop);
    return WN_ERR_INVALID;
  }
  return workspace_layout(const_cast<wn_model*>(m), slice_sz).total;
}

}  // extern "C"
