"""Thin object layer over the C ABI: variable registry + training / generation engines.

PyTorch is used for device buffers and streams only; every computation below is a call into
libwavenet_b200.so (no torch op, no CPU path).
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import WnArch, check, ptr

ARCH_KEYS = ("n_blocks", "n_block_layers", "n_quant", "n_res", "n_dil", "n_skip", "n_post",
             "n_gc_embed", "n_gc_category", "use_bias")
LC_KEYS = ("n_lc_in", "n_lc_out")  # + lc_upsample (list of strides); absent / 0 == no local conditioning


@dataclass
class ParamInfo:
    name: str
    offset: int
    shape: Tuple[int, ...]
    kind: int  # _lib.KIND_FILTER / KIND_BIAS

    @property
    def numel(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n


@dataclass
class SaveInfo:
    name: str  # SAVE_{dil}_{b}_{bl}   (reference arch.py:142, tmodel.py:123)
    offset: int
    dil: int
    shape: Tuple[int, int, int]


class Registry:
    """Host-only mirror of the reference's variable registry (arch.py:85-103,112-167):
    serial names, shapes and the flat-arena layout.  Needs no GPU."""

    def __init__(self, arch: dict, n_slots: int):
        lib = _lib.load()
        self.arch = {k: int(arch[k]) for k in ARCH_KEYS}
        for k in LC_KEYS:
            self.arch[k] = int(arch.get(k, 0) or 0)
        self.arch["lc_upsample"] = [int(x) for x in (arch.get("lc_upsample") or [])] if self.arch["n_lc_out"] > 0 else []
        if self.arch["n_lc_out"] == 0:
            self.arch["n_lc_in"] = 0
        self.lc_hop = int(np.prod(self.arch["lc_upsample"])) if self.arch["lc_upsample"] else 1
        self.n_slots = int(n_slots)
        wa = WnArch.from_dict(self.arch)
        h = C.c_void_p()
        check(lib.wn_model_create(C.byref(wa), self.n_slots, C.byref(h)), "wn_model_create")
        self.handle = h
        self._lib = lib
        self.n_layers = lib.wn_n_layers(h)
        self.recep_field = lib.wn_recep_field(h)
        self.n_param_elems = lib.wn_param_elems(h)
        self.save_elems = lib.wn_save_elems(h)
        self.params: "OrderedDict[str, ParamInfo]" = OrderedDict()
        name = C.create_string_buffer(128)
        off, nd, kind = C.c_int64(), C.c_int32(), C.c_int32()
        shp = (C.c_int64 * 3)()
        for i in range(lib.wn_param_count(h)):
            check(lib.wn_param_info(h, i, name, 128, C.byref(off), C.byref(nd), shp, C.byref(kind)))
            nm = name.value.decode()
            self.params[nm] = ParamInfo(nm, off.value, tuple(int(shp[j]) for j in range(nd.value)), kind.value)
        self.saves: List[SaveInfo] = []
        dil = C.c_int32()
        nbl = self.arch["n_block_layers"]
        for l in range(self.n_layers):
            check(lib.wn_save_info(h, l, C.byref(off), C.byref(dil)))
            b, bl = divmod(l, nbl)
            self.saves.append(SaveInfo("SAVE_%d_%d_%d" % (dil.value, b, bl), off.value, dil.value,
                                       (self.n_slots, dil.value, self.arch["n_res"])))

    def workspace_bytes(self, slice_sz: int) -> int:
        n = self._lib.wn_workspace_bytes(self.handle, int(slice_sz))
        if n < 0:
            check(int(n), "wn_workspace_bytes")
        return int(n)

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self._lib.wn_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise _lib.WaveNetLibError("lb_wavenet_b200 needs a CUDA device: there is no CPU fallback")
    return torch


class TrainEngine:
    """Device state + calls for one data-parallel rank's slots."""

    def __init__(self, arch: dict, n_slots: int, device: str = "cuda"):
        torch = _require_cuda()
        self.torch = torch
        self.reg = Registry(arch, n_slots)
        self.lib = self.reg._lib
        self.device = torch.device(device)
        n = self.reg.n_param_elems
        self.params = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.grads = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.m = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.v = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.save = torch.zeros(self.reg.save_elems, dtype=torch.bfloat16, device=self.device)
        self.stats = torch.zeros(_lib.WN_NSTATS, dtype=torch.float64, device=self.device)
        self.ws = None
        self.ws_T = -1

    # ---- variable access ---------------------------------------------------------------
    def view(self, name: str, arena=None):
        """Tensor view of a trainable variable inside the flat arena."""
        info = self.reg.params[name]
        arena = self.params if arena is None else arena
        return arena[info.offset:info.offset + info.numel].view(info.shape)

    def save_view(self, layer: int):
        s = self.reg.saves[layer]
        n = s.shape[0] * s.shape[1] * s.shape[2]
        return self.save[s.offset:s.offset + n].view(s.shape)

    def load_state(self, state: Dict[str, np.ndarray]) -> None:
        """Load variables by serial name (fp32 params; SAVE_* are rounded to bf16)."""
        torch = self.torch
        for name in self.reg.params:
            if name in state:
                self.view(name).copy_(torch.as_tensor(np.asarray(state[name], np.float32)).to(self.device))
        for l, s in enumerate(self.reg.saves):
            if s.name in state:
                self.save_view(l).copy_(torch.as_tensor(np.asarray(state[s.name], np.float32)).to(self.device))

    def export_state(self) -> "OrderedDict[str, np.ndarray]":
        out: "OrderedDict[str, np.ndarray]" = OrderedDict()
        for name in self.reg.params:
            out[name] = self.view(name).detach().cpu().numpy().copy()
        for l, s in enumerate(self.reg.saves):
            out[s.name] = self.save_view(l).float().cpu().numpy()
        return out

    # ---- compute ------------------------------------------------------------------------
    def _ensure_ws(self, T: int):
        if self.ws_T != T:
            nbytes = self.reg.workspace_bytes(T)
            self.ws = None
            self.ws = self.torch.empty(nbytes, dtype=self.torch.uint8, device=self.device)
            self.ws_T = T

    def forward(self, wav, ids, want_logits: bool = False, mel=None):
        """wav, ids: int32 device tensors [n_slots, T]; mel: float32 device tensor [n_slots, T / hop, n_lc_in] when the
        architecture has local conditioning.  Updates SAVE and stats in place."""
        torch = self.torch
        assert wav.dtype == torch.int32 and ids.dtype == torch.int32 and wav.is_cuda and ids.is_cuda
        assert wav.shape == ids.shape and wav.shape[0] == self.reg.n_slots
        wav, ids = wav.contiguous(), ids.contiguous()
        T = int(wav.shape[1])
        if self.reg.arch["n_lc_out"] > 0:
            if mel is None:
                raise ValueError("this architecture has local conditioning: forward() needs the mel frames")
            want = (self.reg.n_slots, T // self.reg.lc_hop, self.reg.arch["n_lc_in"])
            if T % self.reg.lc_hop or tuple(mel.shape) != want or mel.dtype != torch.float32 or not mel.is_cuda:
                raise ValueError("mel must be a float32 device tensor of shape {} (slice_sz {} / hop {}), got {} {}".format(
                    want, T, self.reg.lc_hop, tuple(mel.shape), mel.dtype))
            mel = mel.contiguous()
        else:
            mel = None
        self._ensure_ws(T)
        logits = None
        if want_logits:
            logits = torch.empty(self.reg.n_slots, T, self.reg.arch["n_quant"], dtype=torch.float32, device=self.device)
        check(self.lib.wn_train_forward(self.reg.handle, ptr(self.params), ptr(self.save), ptr(wav), ptr(ids), ptr(mel),
                                        T, ptr(self.ws), ptr(self.stats), ptr(logits), _lib.cur_stream()),
              "wn_train_forward")
        self._last = (wav, ids, T)
        self._last_mel = mel  # keep the buffer alive until the backward has consumed the conditioning planes
        return logits

    def backward(self):
        wav, ids, T = self._last
        check(self.lib.wn_train_backward(self.reg.handle, ptr(self.params), ptr(wav), ptr(ids), T, ptr(self.ws),
                                         ptr(self.grads), _lib.cur_stream()), "wn_train_backward")

    def backward_phases(self, begin: int, end: int):
        """phases [begin, end) of the backward: 0 = post-net, p = layer L-p, L+1 = PRE / GC (wavenet_b200.h)."""
        wav, ids, T = self._last
        check(self.lib.wn_train_backward_phases(self.reg.handle, ptr(self.params), ptr(wav), ptr(ids), T, ptr(self.ws),
                                                ptr(self.grads), int(begin), int(end), _lib.cur_stream()),
              "wn_train_backward_phases")

    def adam(self, step: int, lr: float, l2_factor: float, n_valid=None, beta1=0.9, beta2=0.999, eps=1e-8):
        """n_valid: device float64 tensor with the GLOBAL valid count (default: this rank's)."""
        nv = self.stats[_lib.STAT_N_VALID:_lib.STAT_N_VALID + 1] if n_valid is None else n_valid
        check(self.lib.wn_adam_step(self.reg.handle, ptr(self.params), ptr(self.grads), ptr(self.m), ptr(self.v),
                                    ptr(nv), int(step), float(lr), float(l2_factor), float(beta1), float(beta2),
                                    float(eps), _lib.cur_stream()), "wn_adam_step")

    def l2_loss(self):
        check(self.lib.wn_l2_loss(self.reg.handle, ptr(self.params), ptr(self.stats), _lib.cur_stream()), "wn_l2_loss")

    def debug_read(self, what: int, layer: int = 0):
        torch = self.torch
        a = self.reg.arch
        ncols = {0: a["n_res"], 1: a["n_dil"], 2: a["n_skip"], 3: a["n_post"], 4: a["n_quant"], 5: a["n_res"],
                 6: a["n_dil"], 7: a["n_res"], 9: 2 * a["n_dil"], 10: 128, 11: 2 * a["n_dil"]}[what]
        out = torch.empty(self.reg.n_slots, self.ws_T, ncols, dtype=torch.float32, device=self.device)
        check(self.lib.wn_debug_read(self.reg.handle, ptr(self.ws), self.ws_T, what, layer, ptr(out),
                                     _lib.cur_stream()), "wn_debug_read")
        return out

    def read_stats(self) -> dict:
        s = self.stats.cpu().numpy()
        return dict(xent_sum=float(s[0]), n_valid=int(round(s[1])), diff_sum=int(round(s[2])), l2=float(s[3]))


class GenEngine:
    """Incremental generator state for n_streams independent streams on one GPU."""

    def __init__(self, arch: dict, n_streams: int, device: str = "cuda"):
        torch = _require_cuda()
        self.torch = torch
        self.reg = Registry(arch, 1)
        self.lib = self.reg._lib
        self.device = torch.device(device)
        self.n_streams = int(n_streams)
        n = self.lib.wn_gen_workspace_bytes(self.reg.handle, self.n_streams)
        if n < 0:
            check(int(n), "wn_gen_workspace_bytes")
        self.gws = torch.zeros(int(n), dtype=torch.uint8, device=self.device)
        self.params = torch.zeros(self.reg.n_param_elems, dtype=torch.float32, device=self.device)
        self.t = 0
        self.reset()

    def view(self, name: str):
        info = self.reg.params[name]
        return self.params[info.offset:info.offset + info.numel].view(info.shape)

    def reset(self):
        check(self.lib.wn_gen_reset(self.reg.handle, ptr(self.gws), self.n_streams, _lib.cur_stream()), "wn_gen_reset")
        self.t = 0

    def load_state(self, state: Dict[str, np.ndarray], gc_ids=None):
        torch = self.torch
        for name in self.reg.params:
            if name in state:
                self.view(name).copy_(torch.as_tensor(np.asarray(state[name], np.float32)).to(self.device))
        self.load_params(self.params, gc_ids)

    def load_params(self, params, gc_ids=None):
        torch = self.torch
        g = None
        if gc_ids is not None:
            g = torch.as_tensor(np.asarray(gc_ids, np.int32)).to(self.device)
        self._gc = g
        check(self.lib.wn_gen_load_params(self.reg.handle, ptr(params), ptr(g), ptr(self.gws), self.n_streams,
                                          _lib.cur_stream()), "wn_gen_load_params")

    def run(self, n_steps: int, seed: int, teacher=None, want_logits: bool = False):
        """Advance every stream n_steps; returns int32 codes [n_streams, n_steps] (device)."""
        torch = self.torch
        out = torch.empty(self.n_streams, n_steps, dtype=torch.int32, device=self.device)
        logits = None
        if want_logits:
            logits = torch.empty(self.n_streams, n_steps, self.reg.arch["n_quant"], dtype=torch.float32,
                                 device=self.device)
        tt, nt = None, 0
        if teacher is not None:
            tt = torch.as_tensor(np.asarray(teacher, np.int32)).to(self.device)
            nt = int(tt.numel())
        check(self.lib.wn_gen_run(self.reg.handle, ptr(self.gws), self.n_streams, self.t, int(n_steps),
                                  int(seed) & (2 ** 64 - 1), ptr(tt), nt, ptr(out), ptr(logits), _lib.cur_stream()),
              "wn_gen_run")
        self.t += int(n_steps)
        return (out, logits) if want_logits else out
