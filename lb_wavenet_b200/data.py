"""Pinned, double-buffered mu-law window loader with the reference's slot-dealing semantics.

Mirrors reference data.py (MaskedSliceWav): a catalog TSV ``voice_id<TAB>wav.npy<TAB>mel.npy``
(data.py:43-48), an endlessly repeated, buffer-shuffled file stream (data.py:246-250), B slot
generators that share that one stream and concatenate files end to end into exact ``slice_sz``
windows with an id / validity mask (data.py:110-227): the first F-1 samples of every file carry
id 0 == invalid (data.py:133,156-159).

Mechanism (new): the dealing is replayed with array slicing instead of per-slice np.append under
the GIL inside a tf.data generator thread; a worker thread fills pinned host buffers ahead of
the consumer and uploads them on a copy stream, so the training stream only waits on an event.
Data-parallel ranks replay the SAME dealing (file -> slot assignment needs file lengths only, read
from the .npy headers) and materialise just their own slots.
"""
from __future__ import annotations

import itertools
import os
import queue
import threading
from dataclasses import dataclass
from sys import maxsize, stderr
from typing import Callable, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np

from . import ckpt

WavSource = Union[str, np.ndarray]


def shuffled_repeat_order(n_files: int, seed: int, skip: int = 0) -> Iterator[int]:
    """ds.repeat().shuffle(buffer_size=n_files, seed).skip(k) (reference data.py:246-250): a
    streaming shuffle buffer over the endlessly repeated catalog.  TensorFlow's own RNG stream is
    not reproducible without TensorFlow; the buffer algorithm is the same, drawn from numpy PCG64."""
    rng = np.random.Generator(np.random.PCG64(seed & ((1 << 63) - 1)))
    buf = list(range(n_files))
    nxt = 0
    produced = 0
    while True:
        j = int(rng.integers(0, len(buf)))
        val = buf[j]
        buf[j] = nxt
        nxt = (nxt + 1) % n_files
        if produced >= skip:
            yield val
        produced += 1


def _npy_len(src: WavSource) -> int:
    if isinstance(src, np.ndarray):
        return int(src.shape[0])
    return int(np.load(src, mmap_mode="r").shape[0])


def _npy_load(src: WavSource) -> np.ndarray:
    return src if isinstance(src, np.ndarray) else np.load(src)


_NP_DTYPE_CODE = {np.dtype(np.uint8): 0, np.dtype(np.int16): 1, np.dtype(np.int32): 2, np.dtype(np.int64): 3,
                  np.dtype(np.float32): 4, np.dtype(np.float64): 5}
_WAV_DTYPES = {"u8": (np.uint8, 0), "i32": (np.int32, 2), "f32": (np.float32, 4)}


class SlotDealer:
    """Deterministic replay of reference data.py:110-227 for slots [slot_lo, slot_hi) of batch_sz.

    Files are pulled lazily from ONE shared stream in slot order, exactly when a slot's generator
    would call next(wav_gen) (data.py:140): which slot receives which file therefore matches the
    reference for any mix of file lengths.

    The cursor arithmetic over all global slots and the copies into the batch buffer are two C calls
    (wn_deal_plan / wn_deal_fill, csrc/dealer.cpp) made through ctypes, i.e. WITHOUT the GIL: the loader thread no
    longer competes with the training thread for the interpreter.  This class keeps the file stream (numpy PCG64
    shuffle buffer), the .npy loading of the local slots' current files and the state for exact resume.

    wav_dtype: 'i32' (mu-law codes as int32, data.py:262-265), 'u8' (the same codes in one byte: what the loader ships
    to the device) or 'f32' (raw float audio for wav_input_type == 'raw', tmodel.py:59-62).
    """

    def __init__(self, catalog: Sequence[Tuple[int, WavSource]], batch_sz: int, slice_sz: int,
                 recep_field_sz: int, mel_hop_sz: int = 1, seed: int = 0, position: int = 0,
                 slot_lo: int = 0, slot_hi: Optional[int] = None, quiet: bool = False, wav_dtype: str = "i32",
                 mel_channels: int = 0):
        """catalog entries: (voice_id, wav) or (voice_id, wav, mel); mel_channels > 0: also deal the mel frames
        [len / hop, mel_channels] of every window (reference data.py:121,170-171,187,222)."""
        if not catalog:
            raise ValueError("empty sample catalog")
        from . import _lib
        self._lib = _lib.load()
        self._check = _lib.check
        self.catalog = list(catalog)
        self.batch_sz, self.slice_sz = int(batch_sz), int(slice_sz)
        self.F, self.hop = int(recep_field_sz), int(mel_hop_sz)
        self.slot_lo, self.slot_hi = slot_lo, batch_sz if slot_hi is None else slot_hi
        self.quiet = quiet
        self.wav_dtype = wav_dtype
        self._np_wav, self._out_code = _WAV_DTYPES[wav_dtype]
        self._order = shuffled_repeat_order(len(self.catalog), seed, position)
        self._order_buf = np.empty(0, np.int32)   # upcoming entries of the shared file stream, not yet consumed
        n, nf = self.batch_sz, len(self.catalog)
        self._datum = np.array([int(position)], np.int64)   # data.py:79
        self._cur_file = np.full(n, -1, np.int64)            # catalog index of the slot's current file
        self._cur_pos = np.zeros(n, np.int64)                # cursor into it
        self._cur_len = np.zeros(n, np.int64)                # usable (hop-trimmed) length
        self._slot_count = np.full(n, int(position), np.int64)  # datum_count of the slot's latest pull
        self._usable = np.empty(nf, np.int64)
        for i in range(nf):
            ln = _npy_len(self.catalog[i][1])
            self._usable[i] = ln - (ln % self.hop)            # data.py:141-142
        if int(self._usable.max()) < self.F:
            raise ValueError("every file is shorter than the receptive field {}".format(self.F))
        self._voice = np.array([int(c[0]) for c in self.catalog], np.int32)
        self._file_ptr = np.zeros(nf, np.uint64)
        self._file_dtype = np.zeros(nf, np.int32)
        self._loaded = {}                                     # catalog index -> contiguous array (local slots' files)
        self.mel_channels = int(mel_channels)
        self._mel_loaded = {}
        if self.mel_channels > 0 and any(len(c) < 3 or c[2] is None for c in self.catalog):
            raise ValueError("local conditioning needs a mel.npy for every catalog entry (data.py:43-48)")
        self._seg = np.empty((max(64, 8 * n), 5), np.int64)
        self._n_seg = np.zeros(1, np.int64)
        self._used = np.zeros(1, np.int64)

    @property
    def _datum_count(self) -> int:
        return int(self._datum[0])

    def _usable_len(self, idx: int) -> int:
        return int(self._usable[idx])

    def _load(self, idx: int) -> None:
        if idx in self._loaded:
            return
        arr = np.ascontiguousarray(np.asarray(_npy_load(self.catalog[idx][1])))
        code = _NP_DTYPE_CODE.get(arr.dtype)
        if code is None:
            arr = np.ascontiguousarray(arr.astype(np.float32 if arr.dtype.kind == "f" else np.int64))
            code = _NP_DTYPE_CODE[arr.dtype]
        self._loaded[idx] = arr
        self._file_ptr[idx] = arr.ctypes.data
        self._file_dtype[idx] = code
        if self.mel_channels > 0:
            mel = np.asarray(_npy_load(self.catalog[idx][2]), np.float32)
            if mel.ndim != 2 or mel.shape[1] != self.mel_channels or mel.shape[0] * self.hop != int(self._usable[idx]):
                # data.py:143-146
                raise ValueError("Error: len(wav) = {}, len(mel) * mel_hop_sz = {} (mel shape {}, expected {} channels)"
                                 .format(int(self._usable[idx]), mel.shape[0] * self.hop, mel.shape, self.mel_channels))
            self._mel_loaded[idx] = mel

    def _evict(self) -> None:
        live = set(int(f) for f in self._cur_file[self.slot_lo:self.slot_hi] if f >= 0)
        for idx in [i for i in self._loaded if i not in live]:
            del self._loaded[idx]
            self._mel_loaded.pop(idx, None)
            self._file_ptr[idx] = 0

    # ---- exact resume (beyond the reference, whose (seed, position) restart re-deals every slot from a fresh file:
    # data.py:249-250,280-286) -------------------------------------------------------------------------------------
    def state(self) -> dict:
        """Everything that determines all future batches, as small int64 arrays: the shared stream's position and, per
        slot, the current file (catalog index, -1 = none yet), the cursor into it and its pull count."""
        return {"stream_position": np.array(self._datum_count, np.int64),
                "slot_file": self._cur_file.copy(), "slot_pos": self._cur_pos.copy(),
                "slot_count": self._slot_count.copy()}

    def load_state(self, st: dict, seed: int) -> None:
        if len(st["slot_file"]) != self.batch_sz:
            raise ValueError("loader state is for batch_sz {}, not {}".format(len(st["slot_file"]), self.batch_sz))
        self._datum[0] = int(st["stream_position"])
        self._order = shuffled_repeat_order(len(self.catalog), seed, self._datum_count)
        self._order_buf = np.empty(0, np.int32)
        self._cur_file[:] = np.asarray(st["slot_file"], np.int64)
        self._cur_pos[:] = np.asarray(st["slot_pos"], np.int64)
        self._slot_count[:] = np.asarray(st["slot_count"], np.int64)
        for slot in range(self.batch_sz):
            idx = int(self._cur_file[slot])
            self._cur_len[slot] = self._usable_len(idx) if idx >= 0 else 0
        self._loaded.clear()
        self._mel_loaded.clear()
        self._file_ptr[:] = 0

    def next_batch(self, wav_out: Optional[np.ndarray] = None, ids_out: Optional[np.ndarray] = None,
                   mel_out: Optional[np.ndarray] = None):
        """Returns (latest_file_read_count, wav[n_local x T], ids[int32 n_local x T]); with mel_channels > 0 the mel
        frames float32 [n_local, T / hop, mel_channels] are written into mel_out (kept in self.last_mel)."""
        T, nl = self.slice_sz, self.slot_hi - self.slot_lo
        wav = np.empty((nl, T), self._np_wav) if wav_out is None else wav_out
        ids = np.empty((nl, T), np.int32) if ids_out is None else ids_out
        if wav.dtype != self._np_wav or ids.dtype != np.int32 or not (wav.flags.c_contiguous and ids.flags.c_contiguous):
            raise ValueError("batch buffers must be C-contiguous {} / int32".format(np.dtype(self._np_wav).name))
        lib = self._lib
        while True:
            rc = lib.wn_deal_plan(self.batch_sz, T, self.slot_lo, self.slot_hi, self.F, self._cur_file.ctypes.data,
                                  self._cur_pos.ctypes.data, self._cur_len.ctypes.data, self._slot_count.ctypes.data,
                                  self._datum.ctypes.data, self._order_buf.ctypes.data, len(self._order_buf),
                                  self._used.ctypes.data, self._usable.ctypes.data, len(self.catalog),
                                  self._seg.ctypes.data, len(self._seg), self._n_seg.ctypes.data)
            if rc == 1:  # the file stream buffer ran dry: draw more of the shuffled order (numpy PCG64, see above)
                more = np.fromiter(itertools.islice(self._order, 2 * self.batch_sz + 64), np.int32)
                self._order_buf = np.concatenate([self._order_buf, more])
                continue
            if rc == 2:
                self._seg = np.empty((2 * len(self._seg), 5), np.int64)
                continue
            self._check(rc, "wn_deal_plan")
            break
        self._order_buf = self._order_buf[int(self._used[0]):].copy()
        seg = self._seg[:int(self._n_seg[0])]
        for row in seg:
            if row[0] < 0:
                if not self.quiet:
                    print("Warning: skipping length {} wav file (voice id {}).  Shorter than receptive "
                          "field size of {}".format(int(row[4]), self.catalog[int(row[2])][0], self.F), file=stderr)
            else:
                self._load(int(row[2]))
        self._check(lib.wn_deal_fill(seg.ctypes.data, len(seg), self._file_ptr.ctypes.data, self._file_dtype.ctypes.data,
                                     self._voice.ctypes.data, len(self.catalog), self.F, T, self._out_code,
                                     wav.ctypes.data, ids.ctypes.data), "wn_deal_fill")
        if self.mel_channels > 0:
            hop = self.hop
            mel = np.empty((nl, T // hop, self.mel_channels), np.float32) if mel_out is None else mel_out
            for row in seg:  # windows, files and cursors are all multiples of hop (data.py:32-37,141-142)
                if row[0] >= 0:
                    r, d0, idx, s0, n = (int(x) for x in row)
                    mel[r, d0 // hop:(d0 + n) // hop] = self._mel_loaded[idx][s0 // hop:(s0 + n) // hop]
            self.last_mel = mel
        self._evict()
        return int(self._slot_count[self.batch_sz - 1]), wav, ids  # data.py:220


@dataclass
class Batch:
    file_read_count: int
    wav: object  # int32 [n_local_slots, T] (device tensor when CUDA is present, else numpy)
    ids: object
    mel: object = None
    state: object = None  # SlotDealer.state() right after this batch was dealt (exact resume)


class BatchField:
    """What get_op() hands out in place of a tf.Tensor: a named field of the dataset's batches."""

    def __init__(self, dataset: "MaskedSliceWav", field: str):
        self.dataset, self.field = dataset, field


class MaskedSliceWav(ckpt.Checkpoint):
    """Same constructor and method surface as reference data.py:21-293 (``sess`` is ignored)."""

    def __init__(self, sess, sam_file, sample_rate, slice_sz, prefetch_sz, mel_spectrum_sz, mel_hop_sz,
                 batch_sz, n_keep_checkpoints, ckpt_path, resume_step, dist=None, device: Optional[str] = None,
                 random_seed: Optional[int] = None, wav_input_type: str = "mu_law_quant"):
        super().__init__(ckpt_path, n_keep_checkpoints, resume_step, sess)
        self.sam_file = sam_file
        self.sample_rate = sample_rate
        self.prefetch_sz = max(1, int(prefetch_sz))
        self.mel_spectrum_sz = mel_spectrum_sz
        self.mel_hop_sz = max(1, int(mel_hop_sz))
        if slice_sz % self.mel_hop_sz != 0:  # data.py:32-37
            requested = slice_sz
            slice_sz += self.mel_hop_sz - (slice_sz % self.mel_hop_sz)
            print("Warning: aligning slice size from {} to {} for mel_hop_sz {}".format(
                requested, slice_sz, self.mel_hop_sz), file=stderr)
        self.slice_sz = slice_sz
        self.batch_sz = batch_sz
        self.random_seed = int(np.random.randint(maxsize)) if random_seed is None else int(random_seed)  # data.py:39
        self.ckpt_position = 0  # data.py:41
        self.dist = dist
        self.device = device
        # 'mu_law_quant': .npy files hold mu-law codes, shipped as uint8 and widened on the device; 'raw': float audio in
        # [-1, 1], shipped as float32 and mu-law encoded on the device by the model (tmodel.py:59-62)
        self.wav_input_type = wav_input_type
        self.sample_catalog: List[list] = []
        self.recep_field_sz = None
        self._worker = None
        self._q: Optional[queue.Queue] = None
        self._stop = threading.Event()

    # ---- catalog ------------------------------------------------------------------------
    def init_sample_catalog(self, entries: Optional[Sequence[Tuple[int, WavSource]]] = None):
        """reference data.py:43-48; ``entries`` lets synthetic in-memory 'files' stand in for a TSV."""
        self.sample_catalog = []
        if entries is not None:
            for e in entries:
                self.sample_catalog.append([int(e[0]), e[1], e[2] if len(e) > 2 else None])
            return
        with open(self.sam_file) as sam_fh:
            for s in sam_fh.readlines():
                if not s.strip():
                    continue
                parts = s.rstrip("\n").split("\t")
                vid, wav_path = parts[0], parts[1]
                mel_path = parts[2] if len(parts) > 2 else None
                self.sample_catalog.append([int(vid), wav_path, mel_path])

    def set_receptive_field_size(self, r_sz):
        self.recep_field_sz = r_sz

    def get_max_id(self):
        return max(self.sample_catalog, key=lambda x: x[0])[0]

    # ---- pipeline ---------------------------------------------------------------------------
    def build(self):
        """reference data.py:230-278: create the (restartable) batch stream and register the two
        saveable scalars that determine where a resumed run continues."""
        if self.recep_field_sz is None:
            raise ValueError("set_receptive_field_size() must be called before build()")
        self.add_saveable_objects({
            "random_seed": ckpt.Variable("random_seed", (), np.int64, lambda: np.array(self.random_seed, np.int64),
                                         lambda v: setattr(self, "random_seed", int(v)), trainable=False),
            "ckpt_position": ckpt.Variable("ckpt_position", (), np.int64,
                                           lambda: np.array(self.ckpt_position, np.int64),
                                           lambda v: setattr(self, "ckpt_position", int(v)), trainable=False),
        })
        # optional extra keys: the dealer's exact state after the last CONSUMED batch (the reference's two scalars
        # restart every slot on a fresh file).  Absent from reference-written checkpoints -> approximate restart.
        self._exact_state = None
        n = self.batch_sz

        def getter(key, shape):
            def get():
                st = self._exact_state
                if st is None:  # nothing consumed yet: "no current file" for every slot, stream at ckpt_position
                    st = {"stream_position": np.array(self.ckpt_position, np.int64),
                          "slot_file": np.full(n, -1, np.int64), "slot_pos": np.zeros(n, np.int64),
                          "slot_count": np.full(n, self.ckpt_position, np.int64)}
                return np.asarray(st[key], np.int64).reshape(shape)
            return get

        def setter(key):
            def set_(v):
                if self._exact_state is None:
                    self._exact_state = {}
                self._exact_state[key] = np.array(v, np.int64)
            return set_

        self.add_saveable_objects({
            key: ckpt.Variable(key, shape, np.int64, getter(key, shape), setter(key), trainable=False, optional=True)
            for key, shape in (("stream_position", ()), ("slot_file", (n,)), ("slot_pos", (n,)), ("slot_count", (n,)))})
        self.add_initializable_ops([self._start])

    def _start(self):
        """(Re)start the stream from (random_seed, ckpt_position) -- the initialisable iterators of
        data.py:254,270."""
        self._shutdown()
        lo, hi = (0, self.batch_sz) if self.dist is None else self.dist.slot_range(self.batch_sz)
        self._n_local = hi - lo
        self._use_cuda = False
        if self.device is None or str(self.device).startswith("cuda"):
            try:
                import torch
                self._use_cuda = torch.cuda.is_available()
            except ImportError:
                self._use_cuda = False
        raw = self.wav_input_type == "raw"
        # device path: codes travel as uint8 (5 bytes per timestep with the int32 id) and are widened to
        # int32 on the copy stream; host path (no CUDA: CPU tests, tools): int32 codes as data.py:262-265
        wav_dtype = "f32" if raw else ("u8" if self._use_cuda else "i32")
        # mel frames travel with the windows when the catalog has them and the model consumes them (mel_spectrum_sz > 0)
        self._mel_ch = int(self.mel_spectrum_sz or 0) if all(
            len(e) > 2 and e[2] is not None and (not isinstance(e[2], str) or os.path.exists(e[2]))
            for e in self.sample_catalog) else 0
        self._dealer = SlotDealer([(e[0], e[1], e[2] if len(e) > 2 else None) for e in self.sample_catalog], self.batch_sz,
                                  self.slice_sz, self.recep_field_sz, self.mel_hop_sz, self.random_seed,
                                  self.ckpt_position, lo, hi, wav_dtype=wav_dtype, mel_channels=self._mel_ch)
        st = getattr(self, "_resume_state", None)
        if st is not None:  # exact resume: restore() found the optional keys
            self._dealer.load_state(st, self.random_seed)
            self._resume_state = None
        self._stop = threading.Event()
        self._q = queue.Queue()
        # ring of prefetch_sz + 2 buffers: one being consumed, prefetch_sz ready, one being filled.
        # A buffer returns to the free list with an event recorded on the consumer's stream; the
        # producer waits for that event before overwriting the pinned / device pair.
        n = self.prefetch_sz + 2
        self._free = queue.Queue()
        for k in range(n):
            self._free.put((k, None))
        self._last_k = None
        self._copy_events = []   # (start, end, bytes) of the newest H2D copies: loader_stats()
        self._deal_seconds, self._deal_batches = 0.0, 0
        shape = (self._n_local, self.slice_sz)
        mshape = (self._n_local, self.slice_sz // self.mel_hop_sz, max(1, self._mel_ch))
        if self._use_cuda:
            import torch
            self._torch = torch
            self._dev = torch.device(self.device or "cuda")
            wt = torch.float32 if raw else torch.uint8
            self._pin = [(torch.empty(shape, dtype=wt).pin_memory(),
                          torch.empty(shape, dtype=torch.int32).pin_memory()) for _ in range(n)]
            # staging copy of the pinned pair + the int32 codes the kernels index with (raw audio: float32, encoded by
            # the model)
            self._stage = [torch.empty(shape, dtype=wt, device=self._dev) for _ in range(n)]
            self._devbuf = [(self._stage[k] if raw else torch.empty(shape, dtype=torch.int32, device=self._dev),
                             torch.empty(shape, dtype=torch.int32, device=self._dev)) for k in range(n)]
            self._copy_stream = torch.cuda.Stream(device=self._dev)
            if self._mel_ch:
                self._pin_mel = [torch.empty(mshape, dtype=torch.float32).pin_memory() for _ in range(n)]
                self._dev_mel = [torch.empty(mshape, dtype=torch.float32, device=self._dev) for _ in range(n)]
        else:
            self._hostbuf = [(np.empty(shape, np.float32 if raw else np.int32), np.empty(shape, np.int32))
                             for _ in range(n)]
            if self._mel_ch:
                self._host_mel = [np.empty(mshape, np.float32) for _ in range(n)]
        self._worker = threading.Thread(target=self._produce, name="wav-loader", daemon=True,
                                        args=(self._stop, self._q, self._free, self._dealer))
        self._worker.start()

    def _produce(self, stop, q, free, dealer):
        """Loader thread.  Its queues / stop flag / dealer are arguments, not attributes: a restart (_start) replaces the
        attributes, and a worker that is still finishing a batch must not touch the new ones."""
        import time
        try:
            while not stop.is_set():
                try:
                    k, released = free.get(timeout=0.1)
                except queue.Empty:
                    continue
                if released is not None:
                    released.synchronize()  # the step that read this buffer has finished
                t0 = time.perf_counter()
                if self._use_cuda:
                    torch = self._torch
                    pw, pi = self._pin[k]
                    cnt, _, _ = dealer.next_batch(pw.numpy(), pi.numpy(), self._pin_mel[k].numpy() if self._mel_ch else None)
                    st = dealer.state()
                    self._deal_seconds += time.perf_counter() - t0
                    self._deal_batches += 1
                    dw, di = self._devbuf[k]
                    with torch.cuda.stream(self._copy_stream):
                        e0 = torch.cuda.Event(enable_timing=True)
                        e1 = torch.cuda.Event(enable_timing=True)
                        e0.record(self._copy_stream)
                        self._stage[k].copy_(pw, non_blocking=True)
                        di.copy_(pi, non_blocking=True)
                        if self._mel_ch:
                            self._dev_mel[k].copy_(self._pin_mel[k], non_blocking=True)
                        e1.record(self._copy_stream)
                        if dw is not self._stage[k]:
                            from . import _lib
                            _lib.check(_lib.load().wn_codes_u8_to_i32(self._stage[k].data_ptr(), dw.data_ptr(), dw.numel(),
                                                                      self._copy_stream.cuda_stream), "wn_codes_u8_to_i32")
                        ev = torch.cuda.Event()
                        ev.record(self._copy_stream)
                    self._copy_events.append((e0, e1, pw.numel() * pw.element_size() + pi.numel() * 4 +
                                              (self._pin_mel[k].numel() * 4 if self._mel_ch else 0)))
                    del self._copy_events[:-64]
                    q.put((cnt, k, ev, st))
                else:
                    hw, hi = self._hostbuf[k]
                    cnt, _, _ = dealer.next_batch(hw, hi, self._host_mel[k] if self._mel_ch else None)
                    self._deal_seconds += time.perf_counter() - t0
                    self._deal_batches += 1
                    q.put((cnt, k, None, dealer.state()))
        except Exception as e:  # surface loader failures in the consumer
            q.put(e)

    def loader_stats(self) -> dict:
        """Measured figures of the loader path: host -> device bytes per batch and per timestep, achieved copy
        bandwidth (CUDA events on the copy stream around the two H2D copies of the newest batches) and the host time
        the dealer needs per batch."""
        out = {"batches": self._deal_batches,
               "deal_ms_per_batch": 1e3 * self._deal_seconds / max(1, self._deal_batches)}
        if self._use_cuda and self._copy_events:
            ms, nbytes = 0.0, 0
            for e0, e1, b in list(self._copy_events):
                if e1.query():
                    ms += e0.elapsed_time(e1)
                    nbytes += b
            if ms > 0:
                per_batch = self._copy_events[-1][2]
                out.update(h2d_bytes_per_batch=per_batch,
                           h2d_bytes_per_timestep=per_batch / float(self._n_local * self.slice_sz),
                           h2d_gbs=nbytes / (ms * 1e-3) / 1e9)
        return out

    def next_batch(self) -> Batch:
        """The next [n_local_slots, slice_sz] batch.  Valid until the following next_batch() call."""
        if self._q is None:
            raise ValueError("init_vars() has not been called")
        item = self._q.get()
        if isinstance(item, Exception):
            raise item
        cnt, k, ev, st = item
        self._exact_state = st
        if self._use_cuda:
            torch = self._torch
            cur = torch.cuda.current_stream(self._dev)
            if self._last_k is not None:  # release the previous buffer once everything queued so far is done
                rel = torch.cuda.Event()
                rel.record(cur)
                self._free.put((self._last_k, rel))
            cur.wait_event(ev)  # this batch's H2D copy
            self._last_k = k
            dw, di = self._devbuf[k]
            return Batch(cnt, dw, di, self._dev_mel[k] if self._mel_ch else None, st)
        if self._last_k is not None:
            self._free.put((self._last_k, None))
        self._last_k = k
        hw, hi = self._hostbuf[k]
        return Batch(cnt, hw, hi, self._host_mel[k] if self._mel_ch else None, st)

    def _shutdown(self):
        if self._worker is not None:
            self._stop.set()
            try:
                while True:
                    self._q.get_nowait()
            except queue.Empty:
                pass
            self._worker.join(timeout=30)
            if self._worker.is_alive():
                raise RuntimeError("the wav-loader thread did not stop within 30 s")
            self._worker = None

    def __del__(self):
        try:
            self._shutdown()
        except Exception:
            pass

    # ---- reference surface --------------------------------------------------------------------
    def get_itr(self):
        return self

    def __iter__(self):
        return self

    def __next__(self):
        b = self.next_batch()
        return b.file_read_count, b.wav, b.mel, b.ids

    def get_op(self):
        """(file_read_count, wav, mel, id_mask) handles, reference data.py:292-293 / train.py:182."""
        return (BatchField(self, "file_read_count"), BatchField(self, "wav"), BatchField(self, "mel"),
                BatchField(self, "ids"))

    def save(self, step, read_count):
        """reference data.py:280-286"""
        self.ckpt_position = int(read_count)
        return super().save(step)

    def restore(self, ckpt_file=None):
        self._exact_state = None
        super().restore(ckpt_file)
        keys = ("stream_position", "slot_file", "slot_pos", "slot_count")
        exact = self._exact_state is not None and all(k in self._exact_state for k in keys)
        self._resume_state = dict(self._exact_state) if exact else None
        if not exact:
            self._exact_state = None
        if self._q is not None:
            self._start()  # re-initialise the iterators from the restored seed / position (or the exact state)
