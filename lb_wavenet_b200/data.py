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


class SlotDealer:
    """Deterministic replay of reference data.py:110-227 for slots [slot_lo, slot_hi) of batch_sz.

    Files are pulled lazily from ONE shared stream in slot order, exactly when a slot's generator
    would call next(wav_gen) (data.py:140): which slot receives which file therefore matches the
    reference for any mix of file lengths.
    """

    def __init__(self, catalog: Sequence[Tuple[int, WavSource]], batch_sz: int, slice_sz: int,
                 recep_field_sz: int, mel_hop_sz: int = 1, seed: int = 0, position: int = 0,
                 slot_lo: int = 0, slot_hi: Optional[int] = None, quiet: bool = False):
        if not catalog:
            raise ValueError("empty sample catalog")
        self.catalog = list(catalog)
        self.batch_sz, self.slice_sz = int(batch_sz), int(slice_sz)
        self.F, self.hop = int(recep_field_sz), int(mel_hop_sz)
        self.slot_lo, self.slot_hi = slot_lo, batch_sz if slot_hi is None else slot_hi
        self.quiet = quiet
        self._order = shuffled_repeat_order(len(self.catalog), seed, position)
        self._datum_count = int(position)  # data.py:79
        self._len_cache = {}
        n = self.batch_sz
        self._cur_file = [-1] * n      # catalog index of the slot's current file
        self._cur_pos = [0] * n        # cursor into it
        self._cur_len = [0] * n        # usable (hop-trimmed) length
        self._cur_data = [None] * n    # loaded array (local slots only)
        self._slot_count = [self._datum_count] * n  # datum_count of the slot's latest pull
        usable = [self._usable_len(i) for i in range(len(self.catalog))]
        if max(usable) < self.F:
            raise ValueError("every file is shorter than the receptive field {}".format(self.F))

    def _usable_len(self, idx: int) -> int:
        if idx not in self._len_cache:
            n = _npy_len(self.catalog[idx][1])
            self._len_cache[idx] = n - (n % self.hop)  # data.py:141-142
        return self._len_cache[idx]

    def _pull(self, slot: int) -> None:
        """next(wav_gen) + the length filter (data.py:140-154)."""
        while True:
            idx = next(self._order)
            self._datum_count += 1  # data.py:82
            n = self._usable_len(idx)
            if n < self.F:
                if not self.quiet:
                    print("Warning: skipping length {} wav file (voice id {}).  Shorter than receptive "
                          "field size of {}".format(n, self.catalog[idx][0], self.F), file=stderr)
                continue
            self._cur_file[slot], self._cur_pos[slot], self._cur_len[slot] = idx, 0, n
            self._slot_count[slot] = self._datum_count
            if self.slot_lo <= slot < self.slot_hi:
                self._cur_data[slot] = np.asarray(_npy_load(self.catalog[idx][1]))[:n]
            return

    # ---- exact resume (beyond the reference, whose (seed, position) restart re-deals every slot from a fresh file:
    # data.py:249-250,280-286) -------------------------------------------------------------------------------------
    def state(self) -> dict:
        """Everything that determines all future batches, as small int64 arrays: the shared stream's position and, per
        slot, the current file (catalog index, -1 = none yet), the cursor into it and its pull count."""
        return {"stream_position": np.array(self._datum_count, np.int64),
                "slot_file": np.array(self._cur_file, np.int64), "slot_pos": np.array(self._cur_pos, np.int64),
                "slot_count": np.array(self._slot_count, np.int64)}

    def load_state(self, st: dict, seed: int) -> None:
        if len(st["slot_file"]) != self.batch_sz:
            raise ValueError("loader state is for batch_sz {}, not {}".format(len(st["slot_file"]), self.batch_sz))
        self._datum_count = int(st["stream_position"])
        self._order = shuffled_repeat_order(len(self.catalog), seed, self._datum_count)
        for slot in range(self.batch_sz):
            idx = int(st["slot_file"][slot])
            self._cur_file[slot], self._cur_pos[slot] = idx, int(st["slot_pos"][slot])
            self._slot_count[slot] = int(st["slot_count"][slot])
            self._cur_len[slot] = self._usable_len(idx) if idx >= 0 else 0
            self._cur_data[slot] = None
            if idx >= 0 and self.slot_lo <= slot < self.slot_hi:
                self._cur_data[slot] = np.asarray(_npy_load(self.catalog[idx][1]))[:self._cur_len[slot]]

    def next_batch(self, wav_out: Optional[np.ndarray] = None, ids_out: Optional[np.ndarray] = None):
        """Returns (latest_file_read_count, wav[int32 n_local x T], ids[int32 n_local x T])."""
        T, nl = self.slice_sz, self.slot_hi - self.slot_lo
        wav = np.empty((nl, T), np.int32) if wav_out is None else wav_out
        ids = np.empty((nl, T), np.int32) if ids_out is None else ids_out
        bound = self.F - 1  # data.py:133
        cur_file, cur_pos, cur_len = self._cur_file, self._cur_pos, self._cur_len
        for slot in range(self.batch_sz):
            local = self.slot_lo <= slot < self.slot_hi
            filled = 0
            if not local:  # another rank's slot: replay the cursor arithmetic only
                while filled < T:
                    if cur_file[slot] < 0 or cur_pos[slot] >= cur_len[slot]:
                        self._pull(slot)
                    take = min(T - filled, cur_len[slot] - cur_pos[slot])
                    cur_pos[slot] += take
                    filled += take
                continue
            while filled < T:
                if self._cur_file[slot] < 0 or self._cur_pos[slot] >= self._cur_len[slot]:
                    self._pull(slot)
                pos = self._cur_pos[slot]
                take = min(T - filled, self._cur_len[slot] - pos)
                if local:
                    r = slot - self.slot_lo
                    wav[r, filled:filled + take] = self._cur_data[slot][pos:pos + take]
                    vid = self.catalog[self._cur_file[slot]][0]
                    seg = ids[r, filled:filled + take]
                    seg[:] = vid
                    nz = bound - pos  # positions < F-1 of the file are invalid (data.py:156-159)
                    if nz > 0:
                        seg[:min(nz, take)] = 0
                self._cur_pos[slot] = pos + take
                filled += take
        return self._slot_count[self.batch_sz - 1], wav, ids  # data.py:220


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
                 random_seed: Optional[int] = None):
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
            for vid, wav in entries:
                self.sample_catalog.append([int(vid), wav, None])
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
        self._dealer = SlotDealer([(e[0], e[1]) for e in self.sample_catalog], self.batch_sz, self.slice_sz,
                                  self.recep_field_sz, self.mel_hop_sz, self.random_seed, self.ckpt_position, lo, hi)
        st = getattr(self, "_resume_state", None)
        if st is not None:  # exact resume: restore() found the optional keys
            self._dealer.load_state(st, self.random_seed)
            self._resume_state = None
        self._n_local = hi - lo
        self._use_cuda = False
        if self.device is None or str(self.device).startswith("cuda"):
            try:
                import torch
                self._use_cuda = torch.cuda.is_available()
            except ImportError:
                self._use_cuda = False
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
        if self._use_cuda:
            import torch
            self._torch = torch
            self._dev = torch.device(self.device or "cuda")
            shape = (self._n_local, self.slice_sz)
            self._pin = [(torch.empty(shape, dtype=torch.int32).pin_memory(),
                          torch.empty(shape, dtype=torch.int32).pin_memory()) for _ in range(n)]
            self._devbuf = [(torch.empty(shape, dtype=torch.int32, device=self._dev),
                             torch.empty(shape, dtype=torch.int32, device=self._dev)) for _ in range(n)]
            self._copy_stream = torch.cuda.Stream(device=self._dev)
        else:
            self._hostbuf = [(np.empty((self._n_local, self.slice_sz), np.int32),
                              np.empty((self._n_local, self.slice_sz), np.int32)) for _ in range(n)]
        # The dealer replays the shared file stream for ALL global slots in Python (only lengths for the slots of other
        # ranks), ~2 ms of interpreter time per batch at 256 slots.  With CPython's default 5 ms switch interval the
        # training thread can wait that long for the GIL while the GPU runs dry; 0.2 ms bounds the wait.
        import sys
        if sys.getswitchinterval() > 2e-4:
            sys.setswitchinterval(2e-4)
        self._worker = threading.Thread(target=self._produce, name="wav-loader", daemon=True)
        self._worker.start()

    def _produce(self):
        try:
            while not self._stop.is_set():
                try:
                    k, released = self._free.get(timeout=0.1)
                except queue.Empty:
                    continue
                if released is not None:
                    released.synchronize()  # the step that read this buffer has finished
                if self._use_cuda:
                    torch = self._torch
                    pw, pi = self._pin[k]
                    cnt, _, _ = self._dealer.next_batch(pw.numpy(), pi.numpy())
                    st = self._dealer.state()
                    dw, di = self._devbuf[k]
                    with torch.cuda.stream(self._copy_stream):
                        dw.copy_(pw, non_blocking=True)
                        di.copy_(pi, non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(self._copy_stream)
                    self._q.put((cnt, k, ev, st))
                else:
                    hw, hi = self._hostbuf[k]
                    cnt, _, _ = self._dealer.next_batch(hw, hi)
                    self._q.put((cnt, k, None, self._dealer.state()))
        except Exception as e:  # surface loader failures in the consumer
            self._q.put(e)

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
            return Batch(cnt, dw, di, None, st)
        if self._last_k is not None:
            self._free.put((self._last_k, None))
        self._last_k = k
        hw, hi = self._hostbuf[k]
        return Batch(cnt, hw, hi, None, st)

    def _shutdown(self):
        if self._worker is not None:
            self._stop.set()
            try:
                while True:
                    self._q.get_nowait()
            except queue.Empty:
                pass
            self._worker.join(timeout=5)
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
