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

arch = config.load_arch(os.path.join(ROOT, "par", "arch_classic_3x10.json"))
B, T = 32, 16384
eng = TrainEngine(arch, B)
rng = np.random.default_rng(0)
wav = torch.as_tensor(rng.integers(0, 256, (B, T)).astype(np.int32)).cuda()
ids = torch.ones(B, T, dtype=torch.int32).cuda()
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
