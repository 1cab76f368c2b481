"""WaveNetGen: host-side mirror of reference imodel.py on top of the persistent generator kernel.

Constructor arguments as used by reference generate.py:61-74.  The reference class is stale
against its own arch.py / ckpt.py (wrong base-ctor argument list imodel.py:25-36,
``self.use_gc`` never set imodel.py:53,113, restore() called with arguments it does not take
generate.py:83) -- the surface is kept, the bit rot is not:
  build_graph() -> (i, waveform, wpos) handles, restore(sess, ckpt), init_buffers(sess),
  gen_sz / gc_ids placeholders, run(feed_dict) in place of sess.run(wave_ops, feed_dict).
Documented deviations (SURVEY quirk ledger): PRE bias is applied; every generated sample is
returned (no dropped trailing chunk); ``chunk_sz`` only sets how many timesteps one kernel launch
advances; sampling is seeded (``seed``), Philox4x32-10 + inverse CDF.
"""
from __future__ import annotations

from sys import stderr
from typing import Optional

import numpy as np

from . import _lib, arch as ar, ckpt


class Placeholder:
    def __init__(self, name):
        self.name = name


class WaveNetGen(ar.WaveNetArch):

    def __init__(self, n_blocks, n_block_layers, n_quant, n_res, n_dil, n_skip, n_post1, n_gc_embed,
                 n_gc_category, use_bias, batch_sz, chunk_sz, teacher_vec, seed: Optional[int] = None,
                 device: str = "cuda", dist=None):
        super().__init__(batch_sz, n_quant, n_res, n_dil, n_skip, n_post1, n_gc_embed, n_gc_category,
                         0, 0, False, 1, None, 0, None)
        self.n_blocks, self.n_block_layers = n_blocks, n_block_layers
        self.use_bias = use_bias
        self.use_gc = n_gc_embed > 0
        self.chunk_sz = max(1, int(chunk_sz))
        self.batch_sz = batch_sz
        self.seed = int(np.random.SeedSequence().entropy % (2 ** 63)) if seed is None else int(seed)
        self.device = device
        self.teacher_vec = teacher_vec
        self.teacher_mu = None
        self.engine = None
        self.gen_sz = Placeholder("gen_sz")
        self.gc_ids = Placeholder("gc_ids") if self.use_gc else None
        from . import config
        self._arch_dict = config.engine_arch(dict(
            n_blocks=n_blocks, n_block_layers=n_block_layers, n_quant=n_quant, n_res=n_res, n_dil=n_dil, n_skip=n_skip,
            n_post=n_post1, n_gc_embed=n_gc_embed, n_gc_category=n_gc_category, use_bias=use_bias))
        # independent streams shard across GPUs with no collective
        self.dist = dist
        self.stream_lo, self.stream_hi = (0, batch_sz) if dist is None else dist.slot_range(batch_sz)
        if teacher_vec is not None:
            print("Teacher vec is {} samples long.".format(np.asarray(teacher_vec).shape[0]), file=stderr)

    def _ensure_engine(self):
        if self.engine is None:
            from .engine import GenEngine
            self.engine = GenEngine(self._arch_dict, self.stream_hi - self.stream_lo, self.device)
        return self.engine

    def _make_variable(self, name, shape, arch, trainable):
        eng = self._ensure_engine()
        torch = eng.torch
        get, set_ = ar.padded_accessors(lambda: eng.view(name), shape)
        return ckpt.Variable(name, shape, np.float32, get, set_, trainable)

    def build_graph(self):
        """reference imodel.py:279-303.  Registers the trainable variables (same serial names as the
        trainer, so a trainer checkpoint restores into the generator) and encodes the teacher."""
        eng = self._ensure_engine()
        for name, info in eng.reg.params.items():
            cat, idx, bias = ar.parse_serial_name(name)   # logical (checkpoint) shape; the arena may be zero-extended
            self.vars[name] = self._make_variable(name, self.var_shape(cat, *idx, get_bias=bias), cat, True)
        self.add_saveable_objects(self.vars)
        if self.teacher_vec is not None:  # imodel.py:44-48: ops.mu_encode(teacher_vec) on the device
            from . import ops
            self.teacher_mu = ops.mu_encode(np.asarray(self.teacher_vec, np.float32), self.n_quant)
        self.graph_built = True
        return ("i", "waveform", "wpos")

    def restore(self, sess=None, ckpt_file=None):
        """generate.py:83 calls restore(sess, ckpt); trainer checkpoints hold extra keys (SAVE_*, counters)
        which the generator ignores."""
        if ckpt_file is None and isinstance(sess, str):
            ckpt_file, sess = sess, None
        tensors = ckpt.read_checkpoint(ckpt_file)
        missing = [k for k in self.vars if k not in tensors]
        if missing:
            raise KeyError("checkpoint {} lacks keys: {}".format(ckpt_file, ", ".join(missing[:8])))
        for k, var in self.vars.items():
            var.assign(tensors[k])

    def init_buffers(self, sess=None):
        """reference imodel.py:274-276: zero the lookback (ring) buffers and the pending input."""
        self._ensure_engine().reset()

    def run(self, feed_dict=None, gen_sz: Optional[int] = None, gc_ids=None, return_codes: bool = False):
        """sess.run(wave_ops, feed_dict) stand-in (generate.py:94-110): returns (n, wav_streams, wpos)
        with wav_streams float32 [batch_sz(local), gen_sz] = mu_decode(sampled codes) (imodel.py:181-182)."""
        from . import ops
        eng = self._ensure_engine()
        feed_dict = feed_dict or {}
        n = int(gen_sz if gen_sz is not None else feed_dict[self.gen_sz])
        if self.use_gc:
            ids = gc_ids if gc_ids is not None else feed_dict.get(self.gc_ids)
            if ids is None:
                raise ValueError("global conditioning needs gc_ids (one voice id per stream)")
            ids = np.asarray(ids, np.int32)
            if ids.shape[0] != self.batch_sz:
                raise ValueError("gc_ids must hold one id per stream ({}), got {}".format(self.batch_sz, ids.shape[0]))
            ids = ids[self.stream_lo:self.stream_hi]
        else:
            ids = None
        eng.load_params(eng.params, ids)
        torch = eng.torch
        chunks = []
        done = 0
        while done < n:
            step = min(self.chunk_sz, n - done)
            # seed offset keeps different ranks' streams on different Philox streams via the stream id;
            # stream ids are local indices, so fold the shard base into the seed
            chunks.append(eng.run(step, self.seed + (self.stream_lo << 40),
                                  teacher=None if self.teacher_mu is None else self.teacher_mu))
            done += step
        codes = torch.cat(chunks, dim=1)
        self.last_codes = codes
        if return_codes:
            return n, codes, 0
        return n, ops.mu_decode(codes, self.n_quant), 0
