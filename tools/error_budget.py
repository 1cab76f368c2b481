#!/usr/bin/env python
"""Where the bf16 error of the gradients comes from (CPU only; developer aid).

    python tools/error_budget.py [arch.json] [slots] [slice]          default: par/arch_classic_3x10.json 2 2048

The oracle's hand-written backward carries the CUDA path's rounding points (`emulate_bf16=True`); every point has a
site name (oracle.ROUND_OFF).  This script switches sites off one at a time -- "what if that tensor were kept in fp32"
-- and all but one -- "that tensor alone in bf16" -- and reports the per-tensor relative L2 error of the gradients
against the fp64 statement of the same step (median / max over the trainable tensors).  Sites: w = operand copies of
the weights, x = the residual stream x_l (and with it the SAVE rows), z = gated outputs, h = post-net activations,
dlog / dpost / dz / dv / dx = the backward's stored or operand tensors.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import wavenet_oracle as O  # noqa: E402
from tests import util  # noqa: E402

arch_file = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "par", "arch_classic_3x10.json")
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
T = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
d = json.load(open(arch_file))
a = O.Arch(d["n_blocks"], d["n_block_layers"], d.get("n_quant", 256), d["n_res"], d["n_dil"], d["n_skip"],
           d.get("n_post", d.get("n_post1")), d.get("n_gc_embed", 0), d.get("n_gc_category", 0), bool(d.get("use_bias", True)))
p = util.scaled_params(a, B, 31) if hasattr(util, "scaled_params") else O.init_params(a, B, seed=31, bias_scale=0.2)
wav, ids = util.synth_batch(B, T, max(1, a.n_gc_category), 32)
w, i = torch.as_tensor(wav).long(), torch.as_tensor(ids).long()
SITES = ["w", "x", "z", "h", "dlog", "dpost", "dz", "dv", "dx"]


def grads(off):
    O.ROUND_OFF = set(off)
    pt, save, _ = O.to_torch_params(a, p, B, torch.float64, requires_grad=False)
    g, info = O.train_backward_manual(a, pt, save, w, i, torch.float64, emulate_bf16=True)
    O.ROUND_OFF = set()
    return {k: v.numpy() for k, v in g.items()}


ref = grads(SITES)  # every rounding off = the fp64 statement


def err(g):
    e = [util.rel_err(g[k], ref[k]) for k in ref if np.abs(ref[k]).max() > 0]
    return float(np.median(e)), float(np.max(e))


print("%s, %d slots x %d timesteps, %d layers: per-tensor relative L2 error of the gradients against fp64 (median / max)"
      % (os.path.basename(arch_file), B, T, a.n_layers))
print("%-34s %8s %8s" % ("rounding sites in bf16", "median", "max"))
print("%-34s %8.4f %8.4f" % ("all (the CUDA path's contract)", *err(grads([]))))
for s in SITES:
    print("%-34s %8.4f %8.4f" % ("all but %s" % s, *err(grads([s]))))
for s in SITES:
    print("%-34s %8.4f %8.4f" % ("only %s" % s, *err(grads([t for t in SITES if t != s]))))
print("%-34s %8.4f %8.4f" % ("all but x, dx (fp32 streams)", *err(grads(["x", "dx"]))))
print("%-34s %8.4f %8.4f" % ("only w (bf16 operands, fp32 rest)", *err(grads([t for t in SITES if t != "w"]))))
