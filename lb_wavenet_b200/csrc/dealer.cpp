// Host-only: the slot dealer of the window loader (reference data.py:110-227) as two plain C functions.
//
// The reference deals B concatenated-file window streams from ONE shared, shuffled, endlessly repeated file stream
// (data.py:211,246-250): slot s pulls the next file exactly when its generator runs dry (data.py:140), so which slot
// gets which file depends on every slot's cursor.  A data-parallel rank therefore has to replay the cursor arithmetic
// of ALL global slots (file lengths only) and materialise just its own.  In Python that replay held the GIL for ~2 ms
// per batch at 256 slots and starved the training thread (end-to-end efficiency 0.87 at 8 GPUs); here it
// is a few microseconds of C, called through ctypes WITHOUT the GIL, so the loader thread never blocks the step.
//
//   wn_deal_plan  cursor arithmetic for one batch of all global slots -> copy segments for the local slots
//   wn_deal_fill  executes the segments: mu-law codes (u8 / i32) or raw float audio into the (pinned) batch buffer,
//                 voice id / 0 into the id mask (first F-1 samples of every file are invalid, data.py:133,156-159)
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "model.h"

using namespace wn;

namespace {
template <typename S, typename O>
bool copy_codes(const S* src, O* dst, int64_t n, bool range_check) {
  bool ok = true;
  for (int64_t i = 0; i < n; ++i) {
    const S v = src[i];
    if (range_check && (v < (S)0 || (int64_t)v > 255)) ok = false;
    dst[i] = (O)v;
  }
  return ok;
}
}  // namespace

extern "C" {

int wn_deal_plan(int32_t batch_sz, int32_t slice_sz, int32_t slot_lo, int32_t slot_hi, int32_t recep_field,
                 int64_t* h_cur_file, int64_t* h_cur_pos, int64_t* h_cur_len, int64_t* h_slot_count,
                 int64_t* h_datum_count, const int32_t* h_order, int64_t n_order, int64_t* h_order_used,
                 const int64_t* h_usable_len, int64_t n_files, int64_t* h_seg, int64_t seg_cap, int64_t* h_n_seg) {
  if (!h_cur_file || !h_cur_pos || !h_cur_len || !h_slot_count || !h_datum_count || !h_order || !h_order_used ||
      !h_usable_len || !h_seg || !h_n_seg || batch_sz < 1 || slice_sz < 1 || slot_lo < 0 || slot_hi > batch_sz ||
      slot_lo > slot_hi || n_files < 1) {
    set_error("wn_deal_plan: invalid argument");
    return WN_ERR_INVALID;
  }
  // work on copies: nothing is committed unless the whole batch could be planned
  std::vector<int64_t> file(h_cur_file, h_cur_file + batch_sz), pos(h_cur_pos, h_cur_pos + batch_sz),
      len(h_cur_len, h_cur_len + batch_sz), cnt(h_slot_count, h_slot_count + batch_sz);
  int64_t datum = *h_datum_count, used = 0, n_seg = 0;
  auto emit = [&](int64_t a, int64_t b, int64_t c, int64_t d, int64_t e) {
    if (n_seg >= seg_cap) return false;
    int64_t* s = h_seg + 5 * n_seg++;
    s[0] = a; s[1] = b; s[2] = c; s[3] = d; s[4] = e;
    return true;
  };
  for (int32_t slot = 0; slot < batch_sz; ++slot) {
    const bool local = slot >= slot_lo && slot < slot_hi;
    int64_t filled = 0;
    while (filled < slice_sz) {
      if (file[slot] < 0 || pos[slot] >= len[slot]) {
        // next(wav_gen) + the length filter (data.py:140-154): files shorter than the receptive field are skipped
        for (;;) {
          if (used >= n_order) return 1;  // the caller extends the order buffer and calls again
          const int32_t idx = h_order[used++];
          if (idx < 0 || idx >= n_files) {
            set_error("wn_deal_plan: file index %d out of range", idx);
            return WN_ERR_INVALID;
          }
          ++datum;  // data.py:82
          const int64_t n = h_usable_len[idx];
          if (n < recep_field) {
            if (!emit(-1, slot, idx, 0, n)) return 2;  // notice: skipped file (the caller prints the reference's warning)
            continue;
          }
          file[slot] = idx; pos[slot] = 0; len[slot] = n; cnt[slot] = datum;
          break;
        }
      }
      const int64_t take = std::min<int64_t>(slice_sz - filled, len[slot] - pos[slot]);
      if (local && !emit(slot - slot_lo, filled, file[slot], pos[slot], take)) return 2;
      pos[slot] += take;
      filled += take;
    }
  }
  memcpy(h_cur_file, file.data(), sizeof(int64_t) * batch_sz);
  memcpy(h_cur_pos, pos.data(), sizeof(int64_t) * batch_sz);
  memcpy(h_cur_len, len.data(), sizeof(int64_t) * batch_sz);
  memcpy(h_slot_count, cnt.data(), sizeof(int64_t) * batch_sz);
  *h_datum_count = datum;
  *h_order_used = used;
  *h_n_seg = n_seg;
  return WN_OK;
}


// dtype codes: 0 = uint8, 1 = int16, 2 = int32, 3 = int64, 4 = float32, 5 = float64
int wn_deal_fill(const int64_t* h_seg, int64_t n_seg, const uint64_t* h_file_ptr, const int32_t* h_file_dtype,
                 const int32_t* h_voice_id, int64_t n_files, int32_t recep_field, int32_t slice_sz, int32_t out_dtype,
                 void* h_wav_out, int32_t* h_ids_out) {
  if (!h_seg || !h_file_ptr || !h_file_dtype || !h_voice_id || !h_wav_out || !h_ids_out ||
      (out_dtype != 0 && out_dtype != 2 && out_dtype != 4)) {
    set_error("wn_deal_fill: invalid argument");
    return WN_ERR_INVALID;
  }
  const int64_t bound = (int64_t)recep_field - 1;  // data.py:133: positions < F-1 of a file carry id 0 == invalid
  for (int64_t k = 0; k < n_seg; ++k) {
    const int64_t* s = h_seg + 5 * k;
    const int64_t row = s[0], dst_off = s[1], idx = s[2], src_off = s[3], n = s[4];
    if (row < 0) continue;  // skipped-file notice
    if (idx < 0 || idx >= n_files || h_file_ptr[idx] == 0 || dst_off < 0 || dst_off + n > slice_sz) {
      set_error("wn_deal_fill: bad segment (file %lld not loaded?)", (long long)idx);
      return WN_ERR_INVALID;
    }
    const void* base = reinterpret_cast<const void*>(h_file_ptr[idx]);
    const int sd = h_file_dtype[idx];
    const int64_t o = row * (int64_t)slice_sz + dst_off;
    bool ok = true;
    if (out_dtype == 4) {  // wav_input_type == 'raw' (tmodel.py:59-62): float audio, mu-law encoded on the device
      float* dst = static_cast<float*>(h_wav_out) + o;
      if (sd == 4) memcpy(dst, static_cast<const float*>(base) + src_off, sizeof(float) * n);
      else if (sd == 5) copy_codes(static_cast<const double*>(base) + src_off, dst, n, false);
      else { set_error("wn_deal_fill: raw wav input needs float .npy audio (file %lld)", (long long)idx); return WN_ERR_INVALID; }
    } else {
      // mu-law codes; u8 transport (5 bytes per timestep with the id) needs every code in [0, 255]
      const bool rc = out_dtype == 0;
#define WN_COPY(ST)                                                                                          \
  (out_dtype == 0 ? copy_codes(static_cast<const ST*>(base) + src_off, static_cast<uint8_t*>(h_wav_out) + o, n, rc) \
                  : copy_codes(static_cast<const ST*>(base) + src_off, static_cast<int32_t*>(h_wav_out) + o, n, rc))
      switch (sd) {
        case 0: ok = WN_COPY(uint8_t); break;
        case 1: ok = WN_COPY(int16_t); break;
        case 2: ok = WN_COPY(int32_t); break;
        case 3: ok = WN_COPY(int64_t); break;
        default:
          set_error("wn_deal_fill: wav_input_type 'mu_law_quant' needs integer mu-law codes, file %lld holds floats "
                    "(use wav_input_type 'raw')", (long long)idx);
          return WN_ERR_INVALID;
      }
#undef WN_COPY
      if (!ok) {
        set_error("wn_deal_fill: file %lld holds a mu-law code outside [0, 255]", (long long)idx);
        return WN_ERR_INVALID;
      }
    }
    int32_t* ids = h_ids_out + o;
    const int32_t vid = h_voice_id[idx];
    const int64_t nz = std::min<int64_t>(std::max<int64_t>(bound - src_off, 0), n);
    for (int64_t i = 0; i < nz; ++i) ids[i] = 0;
    for (int64_t i = nz; i < n; ++i) ids[i] = vid;
  }
  return WN_OK;
}

}  // extern "C"
