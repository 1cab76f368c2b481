/*
 * wavenet_b200.h -- C ABI of the B200-native lb-wavenet hot path (libwavenet_b200.so).
 *
 * The reference (hrbigelow/lb-wavenet) is pure Python on TensorFlow 1.x and has NO FFI /
 * plugin interface: the hot path is reached through Python constructors and TF op calls.
 * Each entry point below therefore cites the reference *Python* interface it replaces
 * (file:line into the reference tree).  INTEGRATION.md shows the ctypes binding a
 * reference maintainer would add.
 *
 * Conventions
 *   - plain C, no torch types; every pointer named d_* is a DEVICE pointer owned by the
 *     caller (PyTorch tensors in our host code), h_* is a HOST pointer;
 *   - one process per GPU; calls are made on the thread that owns the CUDA context and are
 *     asynchronous on `stream` (a cudaStream_t passed as void*);
 *   - return 0 on success, negative wn_status on error, message via wn_last_error();
 *   - no exceptions cross the boundary, no global state besides the model handle and the
 *     thread-local error string.
 *
 * Numerics contract (DESIGN.md "Numerics"): parameters, gradients and Adam slots are fp32;
 * contraction operands and stored activations (x_l, z_l, SAVE state, ring buffers) are
 * bf16; every contraction accumulates in fp32; loss statistics accumulate in fp64.
 */
#ifndef WAVENET_B200_H
#define WAVENET_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WN_ABI_VERSION 2

typedef enum wn_status {
  WN_OK = 0,
  WN_ERR_INVALID = -1,   /* bad argument / unsupported architecture */
  WN_ERR_CUDA = -2,      /* a CUDA runtime / driver call failed      */
  WN_ERR_UNSUPPORTED = -3
} wn_status;

/* Architecture: the keys of par/arch*.json consumed by WaveNetTrain.__init__
 * (reference tmodel.py:8-24) after host-side normalisation (n_post1 -> n_post, defaults). */
typedef struct wn_arch {
  int32_t n_blocks;
  int32_t n_block_layers;
  int32_t n_quant;        /* must be 256 */
  int32_t n_res;          /* multiple of 16, <= 128 */
  int32_t n_dil;          /* multiple of 16, <= 128 */
  int32_t n_skip;         /* multiple of 16, <= 512 */
  int32_t n_post;         /* multiple of 16, <= 512 */
  int32_t n_gc_embed;     /* 0 == no global conditioning */
  int32_t n_gc_category;
  int32_t use_bias;
  /* local conditioning (reference tmodel.py:68-83,156-160; arch.py:75-80,96-97): mel frames [.., n_lc_in] are upsampled
   * by a chain of n_lc_layers transposed convolutions (width == stride == lc_upsample[i]) to one n_lc_out vector per
   * timestep, which every layer projects onto its SIGNAL / GATE pre-activations.  n_lc_out == 0: none.
   * Limits: n_lc_in, n_lc_out <= 128, n_lc_layers <= 8, n_res == n_dil == 32. */
  int32_t n_lc_in;
  int32_t n_lc_out;
  int32_t n_lc_layers;
  int32_t lc_upsample[8];
} wn_arch;

typedef struct wn_model wn_model; /* opaque */

/* kinds reported by wn_param_info */
#define WN_KIND_FILTER 0 /* trainable, L2-regularised (key without 'BIAS', reference tmodel.py:252-254) */
#define WN_KIND_BIAS 1   /* trainable, not regularised */

/* indices into the device statistics block (double[WN_NSTATS]) */
#define WN_STAT_XENT_SUM 0 /* sum of masked cross entropies  (tmodel.py:245) */
#define WN_STAT_N_VALID 1  /* number of valid positions      (tmodel.py:244) */
#define WN_STAT_DIFF_SUM 2 /* sum |argmax(label)-argmax(logit)|*mask (tmodel.py:240-242) */
#define WN_STAT_L2 3       /* sum 0.5*|w|^2 over filters     (tmodel.py:252-258) */
#define WN_NSTATS 4

int32_t wn_abi_version(void);

/* CRC-32C (Castagnoli) of a HOST buffer, running value in / out (0 to start): the checksum of TensorFlow's
 * tensor-bundle checkpoint files written / read by lb_wavenet_b200/tfbundle.py (reference ckpt.py:41,54-62 ->
 * tf.train.Saver) */
uint32_t wn_crc32c(uint32_t crc, const void* h_data, uint64_t n_bytes);
const char* wn_last_error(void);

/* ---- model handle + variable registry -------------------------------------------------
 * Replaces arch.WaveNetArch.__init__ / the shape table / get_variable
 * (reference arch.py:31-103, 112-167).  Host-only: needs no GPU. */
int wn_model_create(const wn_arch* arch, int32_t n_slots, wn_model** out);
void wn_model_destroy(wn_model* m);
int32_t wn_n_layers(const wn_model* m);
/* reference tmodel.py:50-51 get_recep_field_sz */
int32_t wn_recep_field(const wn_model* m);
/* trainable variables live in one flat fp32 arena; entry i has the reference's serial
 * name (arch.py:126,142), an element offset, a shape (<=3 dims) and a kind */
int32_t wn_param_count(const wn_model* m);
int64_t wn_param_elems(const wn_model* m);
int wn_param_info(const wn_model* m, int32_t i, char* name, int32_t name_cap, int64_t* offset,
                  int32_t* ndim, int64_t* shape3, int32_t* kind);
/* D-separation state: SAVE_{dil}_{b}_{bl} [n_slots, dil, n_res] (reference arch.py:82-83,
 * tmodel.py:123-124) kept as bf16 in one arena; layer l lives at element offset *offset */
int64_t wn_save_elems(const wn_model* m);
int wn_save_info(const wn_model* m, int32_t layer, int64_t* offset, int32_t* dil);
/* activation stash + scratch needed by one stage of slice_sz timesteps */
int64_t wn_workspace_bytes(const wn_model* m, int32_t slice_sz);

/* ---- training hot path -----------------------------------------------------------------
 * wn_train_forward replaces the forward half of WaveNetTrain.build (reference
 * tmodel.py:292-328: encode_input_onehot, _preprocess, _dilated_conv x L, _chan_reduce x L,
 * _postprocess, _loss_fcn) including the SAVE assignment (tmodel.py:165).
 *   d_params   fp32 arena (wn_param_elems)
 *   d_save     bf16 SAVE arena (wn_save_elems), read AND updated in place
 *   d_wav      int32 [n_slots, slice_sz] mu-law codes   (data.py:262-265 dtypes)
 *   d_ids      int32 [n_slots, slice_sz] voice id or 0 == invalid
 *   d_mel      fp32 [n_slots, slice_sz / hop, n_lc_in] mel frames of the window, hop = prod(lc_upsample) (data.py:170-171,
 *              222); NULL without local conditioning
 *   d_ws       workspace (wn_workspace_bytes), keeps the stash for wn_train_backward
 *   d_stats    double[WN_NSTATS]; XENT_SUM, N_VALID, DIFF_SUM are OVERWRITTEN
 *   d_logits   optional fp32 [n_slots, slice_sz, n_quant] (NULL: logits never leave the chip) */
int wn_train_forward(wn_model* m, const float* d_params, void* d_save, const int32_t* d_wav,
                     const int32_t* d_ids, const float* d_mel, int32_t slice_sz, void* d_ws, double* d_stats,
                     float* d_logits, void* stream);

/* wn_train_backward replaces Optimizer.compute_gradients (reference tmodel.py:354-358).
 * Writes d(sum of masked cross entropies)/d(param) -- NOT divided by n_valid, no L2 term --
 * into d_grads (fp32 arena, overwritten).  Data-parallel ranks all-reduce d_grads and
 * d_stats[N_VALID]; the 1/n_valid_global scale and the L2 gradient are applied by
 * wn_adam_step, so the result equals the gradient of tmodel.py:261's total loss. */
int wn_train_backward(wn_model* m, const float* d_params, const int32_t* d_wav,
                      const int32_t* d_ids, int32_t slice_sz, void* d_ws, float* d_grads,
                      void* stream);

/* The same computation issued in phases so that data-parallel ranks can overlap the gradient
 * all-reduce with the rest of backward: phase 0 = post-net (+ every SKIP weight gradient);
 * phase p in [1, L] = layer L-p; phase L+1 = PRE gather + global- and local-conditioning gradients.
 * Phases must be issued in ascending order on one stream; [0, L+2) == wn_train_backward. */
int wn_train_backward_phases(wn_model* m, const float* d_params, const int32_t* d_wav,
                             const int32_t* d_ids, int32_t slice_sz, void* d_ws, float* d_grads,
                             int32_t phase_begin, int32_t phase_end, void* stream);

/* wn_adam_step replaces tf.train.AdamOptimizer(lr).apply_gradients (reference
 * train.py:178,186): TF "epsilon-hat" Adam, beta1=.9 beta2=.999 eps=1e-8 by default.
 *   g = d_grads / max(n_valid,1) (0 if n_valid==0, tmodel.py:246-249) + l2_factor * w [filters]
 *   d_n_valid  device double holding the GLOBAL number of valid positions
 *   step       1-based optimiser step t */
int wn_adam_step(wn_model* m, float* d_params, const float* d_grads, float* d_m, float* d_v,
                 const double* d_n_valid, int32_t step, float lr, float l2_factor, float beta1,
                 float beta2, float eps, void* stream);

/* sum 0.5*|w|^2 over filters -> d_stats[WN_STAT_L2]  (reference tmodel.py:250-258) */
int wn_l2_loss(wn_model* m, const float* d_params, double* d_stats, void* stream);

/* test/debug taps into the stash written by the last wn_train_forward / backward:
 * what: 0 = x_l (layer input, [n_slots, T, n_res]), 1 = z_l ([n_slots, T, n_dil]),
 *       2 = h1 ([n_slots,T,n_skip]), 3 = h2 ([n_slots,T,n_post]), 4 = dlogits [.,.,n_quant],
 *       5 = dx_0 (gradient wrt layer-0 input), 6 = dz plane of `layer` [., ., n_dil], 7 = the data-gradient buffer of
 *       parity layer & 1 (dx_l right after layer l's backward) -- readable between wn_train_backward_phases calls,
 *       9 = local-conditioning plane of `layer` [., ., 2 n_dil] (the projections lc_up . [LC_SIGNAL_l | LC_GATE_l]),
 *       10 = upsampled local conditioning [., ., 128], 11 = gradient wrt plane 9 (dv_l, after that layer's backward);
 *       converted to fp32 into d_out */
int wn_debug_read(wn_model* m, const void* d_ws, int32_t slice_sz, int32_t what, int32_t layer,
                  float* d_out, void* stream);

/* ---- incremental generator -------------------------------------------------------------
 * Replaces WaveNetGen.build_graph + init_buffers + the single sess.run of the while_loop
 * (reference imodel.py:214-303, generate.py:78-110).  Ring buffers (one per layer, length
 * dil, replacing the chunk-shifted lookback buffers imodel.py:88-97,199-201) and bf16 weight
 * copies live in a caller-owned workspace. */
int64_t wn_gen_workspace_bytes(const wn_model* m, int32_t n_streams);
/* zero the rings, set every stream's pending input to the all-zero vector (imodel.py:66-70) */
int wn_gen_reset(wn_model* m, void* d_gws, int32_t n_streams, void* stream);
/* refresh the weight copies (+ per-stream global-conditioning projections, imodel.py:53-56,
 * 113-118) from the fp32 arena; d_gc_ids int32 [n_streams] or NULL */
int wn_gen_load_params(wn_model* m, const float* d_params, const int32_t* d_gc_ids, void* d_gws,
                       int32_t n_streams, void* stream);
/* advance every stream n_steps timesteps starting at absolute step t0.
 *   d_teacher  int32 [n_teacher] codes fed back instead of the sample while step < n_teacher
 *              (imodel.py:260-267), or NULL
 *   d_out      int32 [n_streams, n_steps] sampled codes (imodel.py:179-182)
 *   d_logits   optional fp32 [n_streams, n_steps, n_quant] */
int wn_gen_run(wn_model* m, void* d_gws, int32_t n_streams, int64_t t0, int32_t n_steps,
               uint64_t seed, const int32_t* d_teacher, int32_t n_teacher, int32_t* d_out,
               float* d_logits, void* stream);

/* ---- primitive ops ---------------------------------------------------------------------
 * reference ops.py:4-39 (mu_encode / mu_decode, n_quanta == 256), table driven so that the
 * device result equals the float32 numpy twin on every float32 input in [-1, 1]. */
int wn_mu_encode(const float* d_x, int32_t* d_q, int64_t n, void* stream);
int wn_mu_decode(const int32_t* d_q, float* d_x, int64_t n, void* stream);
/* the generator's sampler on caller-provided logits fp32 [n_rows, 256]; row i uses Philox
 * counter (step, stream = i).  Replaces tf.multinomial (reference imodel.py:179). */
int wn_sample_logits(const float* d_logits, int32_t n_rows, uint64_t seed, int64_t step,
                     int32_t* d_out, void* stream);

/* ---- window loader (HOST functions: no GPU, no handle) --------------------------------------
 * The slot dealer of MaskedSliceWav (reference data.py:110-227: _gen_concat_slice_factory / _gen_slice_batch) as plain
 * cursor arithmetic.  Called through ctypes without the GIL from the loader thread.
 * wn_deal_plan advances ALL batch_sz global slots by one slice_sz window: a slot whose file is used up pulls the next
 * entry of h_order (catalog indices of the shared shuffled stream, data.py:246-250) in slot order, skipping files whose
 * usable length is below recep_field (data.py:150-154).  State arrays (int64[batch_sz]) and *h_datum_count are updated in
 * place.  Output: rows of 5 int64 in h_seg -- (local_slot, dst_off, file_idx, src_off, len) for slots in
 * [slot_lo, slot_hi), or (-1, slot, file_idx, 0, usable_len) as a notice that a short file was skipped.
 * Returns 0, or 1 = h_order exhausted / 2 = h_seg too small (then NOTHING was changed: extend and call again). */
int wn_deal_plan(int32_t batch_sz, int32_t slice_sz, int32_t slot_lo, int32_t slot_hi, int32_t recep_field,
                 int64_t* h_cur_file, int64_t* h_cur_pos, int64_t* h_cur_len, int64_t* h_slot_count,
                 int64_t* h_datum_count, const int32_t* h_order, int64_t n_order, int64_t* h_order_used,
                 const int64_t* h_usable_len, int64_t n_files, int64_t* h_seg, int64_t seg_cap, int64_t* h_n_seg);
/* wn_deal_fill executes the segments into one batch buffer [n_local, slice_sz]: h_file_ptr[i] = host address of file
 * i's samples (0 = not loaded), h_file_dtype[i] in {0 u8, 1 i16, 2 i32, 3 i64, 4 f32, 5 f64}; out_dtype 0 = uint8
 * mu-law codes (error if a code is outside [0, 255]), 2 = int32 codes, 4 = float32 raw audio (wav_input_type 'raw',
 * tmodel.py:59-62).  h_ids_out gets the voice id, or 0 for the first recep_field - 1 samples of every file
 * (data.py:133,156-159). */
int wn_deal_fill(const int64_t* h_seg, int64_t n_seg, const uint64_t* h_file_ptr, const int32_t* h_file_dtype,
                 const int32_t* h_voice_id, int64_t n_files, int32_t recep_field, int32_t slice_sz, int32_t out_dtype,
                 void* h_wav_out, int32_t* h_ids_out);
/* device: widen the uint8 codes the loader ships (5 bytes per timestep with the id) to the int32 [n] the training
 * kernels index with */
int wn_codes_u8_to_i32(const uint8_t* d_u8, int32_t* d_i32, int64_t n, void* stream);

/* ---- self tests (device) ---------------------------------------------------------------
 * C[M,N] fp32 = A[M,K] bf16 (row-major) x B[N,K]^T bf16 (row-major) through the
 * TMA + tcgen05.mma + TMEM pipeline used by the training kernels.  swizzle in {32,64,128}. */
int wn_selftest_umma_gemm(const void* d_a, const void* d_b, float* d_c, int32_t M, int32_t N,
                          int32_t K, int32_t swizzle, void* stream);

/* C[M,N] fp32 += A[K,M]^T x B[K,N] (both bf16 row-major, i.e. MN-major operands: the contraction index
 * is the slow one in memory) through the split-K weight-gradient kernel; zero d_c first. */
int wn_selftest_umma_gemm_tn(const void* d_a, const void* d_b, float* d_c, int32_t M, int32_t N,
                             int32_t K, void* stream);
/* the same contraction on CTA pairs (tcgen05 cta_group::2): two CTAs of a cluster share one copy of every B block;
 * M any, N multiple of 32 <= 256, K multiple of 64 */
int wn_selftest_umma_gemm_pair(const void* d_a, const void* d_b, float* d_c, int32_t M, int32_t N,
                               int32_t K, void* stream);

/* Per-category kernel timing with CUDA events recorded on the launch stream around every kernel
 * launch (bench.py's roofline figures).  Categories: 0 prep/embed/SAVE, 1 layer forward, 2 post-net
 * forward + loss, 3 post-net backward, 4 layer backward (fused kernel; gate backward on the HMMA path), 5 layer
 * backward data gradient (HMMA path only), 6 weight gradients, 7 PRE/GC backward, 8 Adam, 9 generator.
 * wn_prof_enable(0) switches the timing off, (1) times every category, (1 << (c + 1)) | ... only the selected
 * categories (an event pair costs ~1 us of stream time per launch).  wn_prof_collect synchronises the device,
 * writes the summed milliseconds and launch counts of the WN_PROF_NCAT categories to HOST arrays
 * and clears the records. */
#define WN_PROF_NCAT 16
int wn_prof_enable(int32_t on);
/* developer aid: d_buf = device buffer of 32 x 2048 int64 (or NULL to switch off); CTA 0 of the persistent layer
 * kernels (layer `layer`, -1 = all) logs (event << 56 | tile << 40 | clock64) per role warp -- see tools/trace_layer.py */
int wn_debug_trace(void* d_buf, int32_t layer);
int wn_prof_collect(double* h_ms, int64_t* h_launches);

/* number of kernels launched by this library since the last call (for bench.py's
 * gpu_launches claim); resets the counter */
int64_t wn_launch_count_reset(void);

#ifdef __cplusplus
}
#endif
#endif /* WAVENET_B200_H */
