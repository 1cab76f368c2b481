"""WaveNetTrain: host-side mirror of reference tmodel.py on top of the sm_100a library.

Same constructor keywords (reference tmodel.py:8-35, train.py:152-163), same variable names /
shapes / checkpoint keys, same loss bookkeeping and progress line (tmodel.py:263-281).  The graph
construction of the reference (`build`) becomes: register the variables, allocate the device
arena + D-separation state + workspace, and return two small op handles; running them is one
call each into libwavenet_b200.so (forward, backward [+ bucketed all-reduce], Adam).
"""
from __future__ import annotations

from sys import stderr
from typing import Optional

import numpy as np

from . import _lib, arch as ar, ckpt, config
from .dist import DistContext, bucket_plan

INT32_MAX = 2 ** 31 - 1


class AdamOptimizer:
    """tf.train.AdamOptimizer(learning_rate) stand-in (reference train.py:178,186): TF defaults
    beta1=0.9, beta2=0.999, epsilon=1e-8, epsilon-hat update (wn_adam_step).  The reference does not checkpoint the
    slots (ckpt.py:41 saves only the model's dict) and a resumed run restarts Adam from zero; here they travel as OPTIONAL
    extra keys with TensorFlow's slot names ('<var>/Adam', '<var>/Adam_1') plus 'optimizer_step', which a
    reference-trained checkpoint simply lacks."""

    def __init__(self, learning_rate, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.learning_rate, self.beta1, self.beta2, self.epsilon = learning_rate, beta1, beta2, epsilon
        self.t = 0

    def apply_gradients(self, grads_and_vars):
        net = getattr(grads_and_vars, "net", None)
        if net is not None:
            net.bind_optimizer(self)
        return ApplyGradsOp(self, grads_and_vars)


class GradsVarsOp:
    def __init__(self, net):
        self.net = net


class ApplyGradsOp:
    def __init__(self, opt: AdamOptimizer, gv: GradsVarsOp):
        self.opt, self.net = opt, gv.net


class LossOp:
    def __init__(self, net):
        self.net = net


class WaveNetTrain(ar.WaveNetArch):

    def __init__(self,
                 # all args from arch.json
                 n_blocks, n_block_layers, n_quant, n_res, n_dil, n_skip, n_post, n_gc_embed, n_gc_category,
                 n_lc_in, n_lc_out, lc_upsample, use_bias, wav_input_type,
                 # args from par.json
                 batch_sz, l2_factor, add_summary, n_keep_checkpoints, ckpt_path, resume_step, n_valid_total,
                 # other arguments
                 sess=None, print_interval=10,
                 # new: data-parallel context (slots are sharded across ranks) and init seed
                 dist: Optional[DistContext] = None, init_seed: Optional[int] = None, device: str = "cuda"):
        super().__init__(batch_sz, n_quant, n_res, n_dil, n_skip, n_post, n_gc_embed, n_gc_category,
                         n_lc_in, n_lc_out, add_summary, n_keep_checkpoints, ckpt_path, resume_step, sess)
        self.n_blocks = n_blocks
        self.n_block_layers = n_block_layers
        self.lc_upsample = lc_upsample
        self.use_bias = use_bias
        self.wav_input_type = wav_input_type
        self.l2_factor = l2_factor
        self.n_valid_total = n_valid_total
        self.print_interval = print_interval
        self.resume_step = resume_step
        self.dist = dist or DistContext()
        self.device = device
        self.init_seed = init_seed
        self.slot_lo, self.slot_hi = self.dist.slot_range(batch_sz)
        self.n_local_slots = self.slot_hi - self.slot_lo
        self._arch_dict = config.engine_arch(dict(
            n_blocks=n_blocks, n_block_layers=n_block_layers, n_quant=n_quant, n_res=n_res, n_dil=n_dil,
            n_skip=n_skip, n_post=n_post, n_gc_embed=n_gc_embed, n_gc_category=n_gc_category,
            n_lc_in=n_lc_in, n_lc_out=n_lc_out, lc_upsample=lc_upsample, use_bias=use_bias))
        self.engine = None
        self._source = None
        self.global_step = 0      # GLOBAL_STEP   (tmodel.py:223-224)
        self.n_valid_cumul = 0    # VALID_SAMPLES (tmodel.py:225-226); int64 here, saturated to int32 on save
        self.last_stats = None
        self._plan = None

    # reference tmodel.py:50-51
    def get_recep_field_sz(self):
        return self.n_blocks * sum([2 ** l for l in range(self.n_block_layers)])

    # ---- variables ----------------------------------------------------------------------
    def _ensure_engine(self):
        if self.engine is None:
            from .engine import TrainEngine
            self.engine = TrainEngine(self._arch_dict, self.n_local_slots, self.device)
            torch = self.engine.torch
            self._comm_stream = torch.cuda.Stream() if self.dist.world > 1 else None
            self._gstats = torch.zeros(_lib.WN_NSTATS, dtype=torch.float64, device=self.engine.device)
            # the step's statistics are final right after the loss kernel (+ their all-reduce): they are copied to
            # pinned host memory on a side stream while the backward runs, so the blocking loss read of
            # sess.run([apply_grads_op, loss_op]) does not leave the GPU idle between steps
            self._host_stats = torch.zeros(_lib.WN_NSTATS, dtype=torch.float64).pin_memory()
            self._rb_stream = self._comm_stream if self._comm_stream is not None else torch.cuda.Stream()
            self._rb_event = torch.cuda.Event()
        return self.engine

    def _make_variable(self, name, shape, arch, trainable):
        eng = self._ensure_engine()
        torch = eng.torch
        if arch == ar.ArchCat.SAVE:
            layer = next(i for i, s in enumerate(eng.reg.saves) if s.name == name)

            R = self.n_res  # logical width; the arena may hold zero-extended rows (config.engine_arch)

            def get_save():
                full = self.dist.all_gather_cat(eng.save_view(layer)[:, :, :R].float().contiguous(), dim=0)  # [B, dil, R]
                return full.cpu().numpy()

            def set_save(v):
                sv = eng.save_view(layer)
                if sv.shape[2] != R:
                    sv.zero_()
                sv[:, :, :R].copy_(torch.as_tensor(v[self.slot_lo:self.slot_hi]).to(eng.device))

            return ckpt.Variable(name, shape, np.float32, get_save, set_save, trainable=False)
        if arch in (ar.ArchCat.GLOBAL_STEP, ar.ArchCat.VALID_SAMPLES):
            attr = "global_step" if arch == ar.ArchCat.GLOBAL_STEP else "n_valid_cumul"
            return ckpt.Variable(name, (), np.int32,
                                 lambda: np.array(min(getattr(self, attr), INT32_MAX), np.int32),
                                 lambda v: setattr(self, attr, int(v)), trainable=False)
        if name not in eng.reg.params:
            raise KeyError("variable {} is not part of this architecture".format(name))
        get, set_ = ar.padded_accessors(lambda: eng.view(name), shape)
        return ckpt.Variable(name, shape, np.float32, get, set_, trainable=True)

    def bind_optimizer(self, opt: "AdamOptimizer"):
        """Adds the optimiser state to the checkpoint as optional keys (a checkpoint without them restores with zero
        slots and t = 0, which is what the reference does on every resume)."""
        eng = self._ensure_engine()
        torch = eng.torch
        self._optimizer = opt
        extra = {}
        for name, info in eng.reg.params.items():
            cat, idx, bias = ar.parse_serial_name(name)
            shape = tuple(self.var_shape(cat, *idx, get_bias=bias))   # logical (checkpoint) shape
            for arena, suffix in ((eng.m, "/Adam"), (eng.v, "/Adam_1")):
                view = arena[info.offset:info.offset + info.numel].view(info.shape)
                get, set_ = ar.padded_accessors(lambda v=view: v, shape)
                extra[name + suffix] = ckpt.Variable(name + suffix, shape, np.float32, get, set_, trainable=False,
                                                     optional=True)
        extra["optimizer_step"] = ckpt.Variable("optimizer_step", (), np.int64, lambda: np.array(opt.t, np.int64),
                                                lambda x: setattr(opt, "t", int(x)), trainable=False, optional=True)
        self.add_saveable_objects(extra)

    def _register_variables(self):
        """Same get_variable call sequence as the reference graph construction (tmodel.py:292-328)."""
        if self.has_global_cond():
            self.get_variable(ar.ArchCat.GC_EMBED)
        self.get_variable(ar.ArchCat.PRE)
        if self.use_bias:
            self.get_variable(ar.ArchCat.PRE, get_bias=True)
        if self.use_lc_input():  # tmodel.py:68-83 (_preprocess_lc, called right after _preprocess: tmodel.py:307-311)
            for i in range(len(self.lc_upsample)):
                self.get_variable(ar.ArchCat.LC_UPSAMPLE, i)
        for b in range(self.n_blocks):
            for bl in range(self.n_block_layers):
                dil = 2 ** bl
                self.get_variable(ar.ArchCat.SAVE, dil, b, bl, trainable=False)
                for a in (ar.ArchCat.SIGNAL, ar.ArchCat.GATE):
                    self.get_variable(a, b, bl)
                    if self.use_bias:
                        self.get_variable(a, b, bl, get_bias=True)
                if self.has_global_cond():
                    self.get_variable(ar.ArchCat.GC_SIGNAL, b, bl)
                    self.get_variable(ar.ArchCat.GC_GATE, b, bl)
                if self.use_lc_input():  # tmodel.py:156-160
                    self.get_variable(ar.ArchCat.LC_SIGNAL, b, bl)
                    self.get_variable(ar.ArchCat.LC_GATE, b, bl)
                for a in (ar.ArchCat.RESIDUAL, ar.ArchCat.SKIP):
                    self.get_variable(a, b, bl)
                    if self.use_bias:
                        self.get_variable(a, b, bl, get_bias=True)
        for a in (ar.ArchCat.POST1, ar.ArchCat.POST2):
            self.get_variable(a)
            if self.use_bias:
                self.get_variable(a, get_bias=True)
        self.get_variable(ar.ArchCat.GLOBAL_STEP, trainable=False)
        self.get_variable(ar.ArchCat.VALID_SAMPLES, trainable=False)

    def build(self, wav_input=None, lc_input=None, id_mask=None):
        """Registers the model's variables, binds the data source and returns (grads_vars, loss)
        op handles (reference tmodel.py:292-340)."""
        self._source = getattr(wav_input, "dataset", None)
        self._register_variables()
        self.add_saveable_objects(self.vars)  # tmodel.py:330
        self.add_initializable_ops([self._init_variables])
        reg = self.engine.reg
        self._plan = bucket_plan([(n, i.offset, i.numel) for n, i in reg.params.items()], reg.n_layers,
                                 self.n_block_layers, reg.n_param_elems, self.has_global_cond())
        return GradsVarsOp(self), LossOp(self)

    def _init_variables(self):
        """Xavier-uniform filters, zero biases; SAVE gets the same Xavier noise the reference's default
        initialiser gives it (arch.py:63-64,125-134; tmodel.py:123-124); counters zero."""
        eng = self._ensure_engine()
        torch = eng.torch
        seed = self.init_seed if self.init_seed is not None else int(np.random.SeedSequence().entropy % (2 ** 32))
        rng = np.random.default_rng(seed)  # identical on every rank when init_seed is given
        if self.dist.world > 1 and self.init_seed is None:
            raise ValueError("data-parallel training needs an explicit init_seed so that replicas agree")
        eng.params.zero_()
        eng.save.zero_()
        for name, info in eng.reg.params.items():
            if info.kind == _lib.KIND_FILTER:   # Xavier bounds from the LOGICAL shape; padding stays zero
                cat, idx, bias = ar.parse_serial_name(name)
                shape = tuple(self.var_shape(cat, *idx, get_bias=bias))
                sl = tuple(slice(0, d) for d in shape)
                eng.view(name)[sl].copy_(torch.as_tensor(ar.xavier_uniform(shape, rng)).to(eng.device))
        for l, s in enumerate(eng.reg.saves):
            full = ar.xavier_uniform((self.batch_sz, s.dil, self.n_res), rng)
            eng.save_view(l)[:, :, :self.n_res].copy_(torch.as_tensor(full[self.slot_lo:self.slot_hi]).to(eng.device))
        eng.m.zero_()
        eng.v.zero_()
        self.global_step = 0
        self.n_valid_cumul = 0

    def _allreduce_mode(self) -> str:
        """'buckets': gradient all-reduce in a few buckets on a side stream, overlapped with the phased backward;
        'single': one all-reduce of the whole arena after the backward.  WN_ALLREDUCE overrides; by default the
        overlap is used only when the arena is large enough for its transfer time to matter (> 8 MB: the wide
        stack's 25 MB)."""
        mode = getattr(self, "_ar_mode", None)
        if mode is None:
            import os
            mode = os.environ.get("WN_ALLREDUCE", "auto")
            if mode not in ("buckets", "single"):
                mode = "single" if self.engine.reg.n_param_elems * 4 <= (8 << 20) else "buckets"
            self._ar_mode = mode
        return mode

    # ---- one training step ------------------------------------------------------------------
    def _prepare_mel(self, mel):
        """mel frames [batch_sz or local slots, slice_sz / hop, n_lc_in] -> float32 device tensor of the local slots"""
        if not self.use_lc_input():
            return None
        if mel is None:
            raise ValueError("this architecture has local conditioning (n_lc_out > 0): the step needs the mel frames")
        eng = self.engine
        torch = eng.torch
        if not torch.is_tensor(mel):
            mel = torch.as_tensor(np.asarray(mel, np.float32))
        if mel.shape[0] == self.batch_sz and self.batch_sz != self.n_local_slots:
            mel = mel[self.slot_lo:self.slot_hi]
        return mel.to(eng.device, dtype=torch.float32, non_blocking=True)

    def _prepare_inputs(self, wav, ids):
        eng = self.engine
        torch = eng.torch
        if not torch.is_tensor(wav):
            wav = torch.as_tensor(np.asarray(wav))
        if not torch.is_tensor(ids):
            ids = torch.as_tensor(np.asarray(ids))
        if wav.shape[0] == self.batch_sz and self.batch_sz != self.n_local_slots:
            wav, ids = wav[self.slot_lo:self.slot_hi], ids[self.slot_lo:self.slot_hi]
        if not wav.is_cuda:
            wav = wav.to(eng.device, non_blocking=True)
        if not ids.is_cuda:
            ids = ids.to(eng.device, non_blocking=True)
        if self.wav_input_type == "raw" or wav.dtype.is_floating_point:  # tmodel.py:59-62
            codes = torch.empty(wav.shape, dtype=torch.int32, device=eng.device)
            x = wav.to(torch.float32).contiguous()
            _lib.check(eng.lib.wn_mu_encode(x.data_ptr(), codes.data_ptr(), x.numel(), _lib.cur_stream()), "wn_mu_encode")
            wav = codes
        return wav.to(torch.int32), ids.to(torch.int32)

    def forward_backward(self, wav, ids, want_loss: bool = False, mel=None):
        """Forward + backward (+ data-parallel reduction).  Leaves the global statistics in
        self._gstats and the summed unnormalised gradients in engine.grads.  want_loss: also evaluate the L2 term (on the
        weights the loss is computed with, as tmodel.py:250-261 does) and start the device -> host copy of the
        statistics as soon as they are final; _finish_step waits for it."""
        eng = self._ensure_engine()
        torch = eng.torch
        wav, ids = self._prepare_inputs(wav, ids)
        eng.forward(wav, ids, mel=self._prepare_mel(mel))
        if want_loss:
            eng.l2_loss()
        self._gstats.copy_(eng.stats)
        L = eng.reg.n_layers
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        if self.dist.world > 1 and self._allreduce_mode() == "single":
            # small models (the 3x10 stack's arena is 2.2 MB): ONE all-reduce after the backward.  Over NVSwitch it costs
            # a few tens of microseconds, less than what bucketed all-reduces running beside the backward take away from
            # the persistent layer kernels (their grids own every SM; measured r1: +4 us per layer launch at 8 GPUs).
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(ev)
                self.dist.all_reduce_sum_(self._gstats[:3])
                if want_loss:
                    self._host_stats.copy_(self._gstats, non_blocking=True)
                    self._rb_event.record(self._comm_stream)
            eng.backward()
            self.dist.all_reduce_sum_(eng.grads)
            cur.wait_stream(self._comm_stream)
            return
        if self.dist.world == 1:
            if want_loss:
                with torch.cuda.stream(self._rb_stream):
                    self._rb_stream.wait_event(ev)
                    self._host_stats.copy_(self._gstats, non_blocking=True)
                    self._rb_event.record(self._rb_stream)
            eng.backward()
            return
        with torch.cuda.stream(self._comm_stream):
            self._comm_stream.wait_event(ev)
            self.dist.all_reduce_sum_(self._gstats[:3])  # xent_sum, n_valid, diff_sum (tmodel.py:244-249)
            if want_loss:
                self._host_stats.copy_(self._gstats, non_blocking=True)
                self._rb_event.record(self._comm_stream)
        phase = 0
        for phase_end, ranges in self._plan:
            _lib.check(eng.lib.wn_train_backward_phases(
                eng.reg.handle, eng.params.data_ptr(), wav.data_ptr(), ids.data_ptr(), int(wav.shape[1]),
                eng.ws.data_ptr(), eng.grads.data_ptr(), phase, phase_end, _lib.cur_stream()),
                "wn_train_backward_phases")
            phase = phase_end
            ev = torch.cuda.Event()
            ev.record(cur)
            with torch.cuda.stream(self._comm_stream):
                self._comm_stream.wait_event(ev)
                for lo, hi in ranges:
                    if hi > lo:
                        self.dist.all_reduce_sum_(eng.grads[lo:hi])
        assert phase == L + 2
        cur.wait_stream(self._comm_stream)

    def train_step(self, wav, ids, optimizer: AdamOptimizer, want_loss: bool = True, mel=None):
        """One optimiser step on a [batch_sz or local slots, slice_sz] batch (host or device tensors).
        Returns the total loss of tmodel.py:261 as a Python float (one 32-byte device->host read)."""
        eng = self._ensure_engine()
        if getattr(self, "_optimizer", None) is not optimizer:
            self.bind_optimizer(optimizer)  # its slots become (optional) checkpoint keys
        self.forward_backward(wav, ids, want_loss, mel=mel)
        optimizer.t += 1
        eng.adam(optimizer.t, optimizer.learning_rate, self.l2_factor, n_valid=self._gstats[1:2],
                 beta1=optimizer.beta1, beta2=optimizer.beta2, eps=optimizer.epsilon)
        if not want_loss:
            self.global_step += 1
            return None
        return self._finish_step(wav)

    def _finish_step(self, wav):
        self._rb_event.synchronize()  # device -> host read of the step's result (copy enqueued by forward_backward)
        s = self._host_stats.numpy().copy()
        xent_sum, n_valid, diff_sum, l2 = float(s[0]), int(round(s[1])), int(round(s[2])), float(s[3])
        mean_xent = xent_sum / n_valid if n_valid != 0 else 0.0  # tmodel.py:246-249
        total = mean_xent + self.l2_factor * l2  # tmodel.py:261
        T = int(wav.shape[1])
        avg_diff = diff_sum // (self.batch_sz * (T - 1)) if T > 1 else 0  # tmodel.py:242: int32 reduce_mean
        self.last_stats = dict(total_loss=total, xent=mean_xent, l2=l2, avg_diff=avg_diff, n_valid=n_valid)
        if self.print_interval and self.global_step % self.print_interval == 0 and self.dist.rank == 0:
            pct = float(self.n_valid_cumul) * 100 / self.n_valid_total if self.n_valid_total else 0.0
            print(('{:5d}\t{:8.4f}\t{:8.4f}\t{:7.2f}\t{:5.0f}\t{:5.0f}\t{:10d}\t{:14d}\t{:5.2f}').format(
                self.global_step, total, mean_xent, l2, avg_diff, n_valid, self.n_valid_cumul,
                self.n_valid_total, pct), file=stderr)  # tmodel.py:265-267
        self.global_step += 1  # tmodel.py:282-284
        self.n_valid_cumul += n_valid
        return total

    def run(self, fetches):
        """sess.run([apply_grads_op, loss_op]) stand-in (reference train.py:240): pulls the next batch
        from the bound dataset and executes one step."""
        if self._source is None:
            raise ValueError("build() was not given dataset ops; use train_step(wav, ids, optimizer)")
        apply_op = next((f for f in fetches if isinstance(f, ApplyGradsOp)), None)
        if apply_op is None:
            raise ValueError("run() needs the op returned by optimizer.apply_gradients")
        batch = self._source.next_batch()
        self.file_read_count = batch.file_read_count
        loss = self.train_step(batch.wav, batch.ids, apply_op.opt, mel=batch.mel)
        return [None if isinstance(f, ApplyGradsOp) else loss for f in fetches]
