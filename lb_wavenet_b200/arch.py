"""Variable registry with the reference's category names, shapes and serial-name grammar.

Host-side mirror of reference arch.py (ArchCat arch.py:6-22, shape table arch.py:85-103,
get_variable arch.py:112-167).  The variables themselves live in the flat device arena laid out
by the C library (wn_param_info); get_variable hands out named views that know how to read and
write their slice, and records them in ``self.vars`` -- the dict the checkpoint is keyed by
(arch.py:142,163; tmodel.py:330).
"""
from __future__ import annotations

from enum import IntEnum
from sys import stderr

import numpy as np

from . import ckpt


class ArchCat(IntEnum):
    # category names are part of the checkpoint-key contract (arch.py:126,142)
    PRE = 1
    LC_UPSAMPLE = 2
    RESIDUAL = 3
    SKIP = 4
    SIGNAL = 5
    GATE = 6
    GC_SIGNAL = 7
    GC_GATE = 8
    GC_EMBED = 9
    LC_SIGNAL = 10
    LC_GATE = 11
    POST1 = 12
    POST2 = 13
    SAVE = 14
    GLOBAL_STEP = 15
    VALID_SAMPLES = 16


def serial_name(arch: ArchCat, *var_indices, get_bias: bool = False) -> str:
    """'_'.join([NAME(+'_BIAS'), *indices])  (reference arch.py:125-126,142)"""
    name = arch.name + ("_BIAS" if get_bias else "")
    return "_".join(map(str, [name, *var_indices]))


def parse_serial_name(name: str):
    """Inverse of serial_name: 'SIGNAL_BIAS_0_3' -> (ArchCat.SIGNAL, (0, 3), True)."""
    for cat in sorted(ArchCat, key=lambda c: -len(c.name)):
        if name == cat.name or name.startswith(cat.name + "_"):
            rest = [x for x in name[len(cat.name):].split("_") if x]
            bias = bool(rest) and rest[0] == "BIAS"
            if bias:
                rest = rest[1:]
            if all(x.isdigit() for x in rest):
                return cat, tuple(int(x) for x in rest), bias
    raise KeyError("not a serial variable name: {}".format(name))


def padded_accessors(view_fn, shape):
    """(get, set) for a variable of logical `shape` stored zero-extended in a (possibly larger) device tensor returned
    by view_fn(): reads the logical slice, writes it and zeroes the padding (config.engine_arch)."""
    sl = tuple(slice(0, int(d)) for d in shape)

    def get():
        return view_fn()[sl].detach().float().cpu().numpy()

    def set_(v):
        import torch
        t = view_fn()
        if tuple(t.shape) != tuple(shape):
            t.zero_()
        t[sl].copy_(torch.as_tensor(np.asarray(v, np.float32)).to(t.device))

    return get, set_


def xavier_uniform(shape, rng: np.random.Generator) -> np.ndarray:
    """tf.contrib.layers.xavier_initializer_conv2d (reference arch.py:63): U(+-sqrt(6/(fan_in+fan_out)))
    with fan_in = shape[-2]*prod(shape[:-2]), fan_out = shape[-1]*prod(shape[:-2])."""
    recept = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
    fan_in, fan_out = shape[-2] * recept, shape[-1] * recept
    bound = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-bound, bound, size=shape).astype(np.float32)


class WaveNetArch(ckpt.Checkpoint):
    """Provides all parameters needed to fully determine a model and manages its variables'
    saving and restoring (same constructor signature as reference arch.py:31-46)."""

    def __init__(self, batch_sz, n_quant, n_res, n_dil, n_skip, n_post, n_gc_embed, n_gc_category,
                 n_lc_in, n_lc_out, add_summary, n_keep_checkpoints, ckpt_path, resume_step, sess=None):
        super().__init__(ckpt_path, n_keep_checkpoints, resume_step, sess)
        self.batch_sz = batch_sz
        self.n_quant = n_quant
        self.n_res = n_res
        self.n_dil = n_dil
        self.n_skip = n_skip
        self.n_post = n_post
        self.n_gc_embed = n_gc_embed
        self.n_gc_category = n_gc_category
        self.n_lc_in = n_lc_in
        self.n_lc_out = n_lc_out
        self.add_summary = add_summary
        self.sess = sess
        self.vars = {}  # serial name -> ckpt.Variable

        def _save_var_shape(dilation, *ignored):
            return [self.batch_sz, dilation, self.n_res]

        def _upsample_shape(i):
            return [self.lc_upsample[i], self.n_lc_out, self.n_lc_in if i == 0 else self.n_lc_out]

        self.shape = {  # reference arch.py:85-103
            ArchCat.PRE: [n_quant, n_res],
            ArchCat.LC_UPSAMPLE: _upsample_shape,
            ArchCat.RESIDUAL: [n_dil, n_res],
            ArchCat.SKIP: [n_dil, n_skip],
            ArchCat.SIGNAL: [2, n_res, n_dil],
            ArchCat.GATE: [2, n_res, n_dil],
            ArchCat.GC_SIGNAL: [n_gc_embed, n_dil],
            ArchCat.GC_GATE: [n_gc_embed, n_dil],
            ArchCat.GC_EMBED: [n_gc_category + 1, n_gc_embed],
            ArchCat.LC_SIGNAL: [n_lc_out, n_dil],
            ArchCat.LC_GATE: [n_lc_out, n_dil],
            ArchCat.POST1: [n_skip, n_post],
            ArchCat.POST2: [n_post, n_quant],
            ArchCat.SAVE: _save_var_shape,
            ArchCat.GLOBAL_STEP: [],
            ArchCat.VALID_SAMPLES: [],
        }

    def has_global_cond(self):
        return self.n_gc_embed > 0

    def use_lc_input(self):
        return self.n_lc_out > 0

    def var_shape(self, arch: ArchCat, *var_indices, get_bias: bool = False):
        shape = self.shape[arch]
        if not isinstance(shape, list):
            shape = shape(*var_indices)
        return [shape[-1]] if get_bias else list(shape)

    def _make_variable(self, name: str, shape, arch: ArchCat, trainable: bool) -> ckpt.Variable:
        """Bind a serial name to storage; overridden by the training / generation models."""
        raise NotImplementedError

    def get_variable(self, arch: ArchCat, *var_indices, get_bias: bool = False, trainable: bool = True,
                     **var_opts) -> ckpt.Variable:
        """Same contract as reference arch.py:112-167: associates a category (+ indices) with a
        shape, registers the variable under its serial name and returns it."""
        shape = self.var_shape(arch, *var_indices, get_bias=get_bias)
        name = serial_name(arch, *var_indices, get_bias=get_bias)
        if name in self.vars:
            var = self.vars[name]
            if list(var.shape) != list(shape):
                print("Attempting to store variable of shape {} under serial name {}.\n"
                      "A variable of shape {} is already stored there.".format(shape, name, var.shape),
                      file=stderr)
                raise SystemExit(1)
            return var
        var = self._make_variable(name, shape, arch, trainable)
        self.vars[name] = var
        return var
