#!/usr/bin/env python
"""Host enqueue time vs device time of one training step at BASELINE configs[1] (developer aid, GPU box only)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lb_wavenet_b200 import config  # noqa: E402
from lb_wavenet_b200.engine import TrainEngine  # noqa: E402

# usage: host_overhead.py [arch.json] [slots] [slice_sz]
arch = config.load_arch(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "par", "arch_classic_3x10.json"))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
T = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
eng = TrainEngine(arch, B)
rng = np.random.default_rng(0)
wav = torch.as_tensor(rng.integers(0, 256, (B, T)).astype(np.int32)).cuda()
ids = torch.as_tensor(rng.integers(1, max(2, arch["n_gc_category"]), (B, 1)).astype(np.int32)).cuda().expand(B, T).contiguous()
print("arch", os.path.basename(sys.argv[1]) if len(sys.argv) > 1 else "classic_3x10", "slots", B, "slice_sz", T)
for _ in range(3):
    eng.forward(wav, ids)
    eng.backward()
    eng.adam(1, 1e-3, 0.0)
torch.cuda.synchronize()
for rep in range(3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t = [time.perf_counter()]
    ev[0].record()
    eng.forward(wav, ids)
    t.append(time.perf_counter())
    ev[1].record()
    eng.backward()
    t.append(time.perf_counter())
    ev[2].record()
    eng.adam(1, 1e-3, 0.0)
    t.append(time.perf_counter())
    ev[3].record()
    torch.cuda.synchronize()
    print("host enqueue ms: fwd %.3f bwd %.3f adam %.3f | device ms (from idle): fwd %.3f bwd %.3f adam %.3f" % (
        1e3 * (t[1] - t[0]), 1e3 * (t[2] - t[1]), 1e3 * (t[3] - t[2]),
        ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])))
    if rep == 2:
        tot = ev[0].elapsed_time(ev[3])
        print("step %.3f ms -> %.2f M output timesteps/s" % (tot, B * (T - 1) / tot / 1e3))

# per-category CUDA-event timing of one more step (ms)
import ctypes as C
from lb_wavenet_b200 import _lib
lib = _lib.load()
cats = ["prep_embed_save", "layer_fwd", "post_fwd_loss", "post_bwd", "layer_bwd", "layer_bwd_data", "wgrad", "pre_gc_bwd",
        "adam", "gen"]
lib.wn_prof_enable(1)
eng.forward(wav, ids)
eng.backward()
eng.adam(1, 1e-3, 0.0)
ms = (C.c_double * 16)()
n = (C.c_int64 * 16)()
lib.wn_prof_collect(ms, n)
lib.wn_prof_enable(0)
print("category ms:", {c: round(ms[k], 3) for k, c in enumerate(cats) if n[k] > 0})
