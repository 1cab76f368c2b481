#!/usr/bin/env python
"""In-kernel timeline of the generator kernel k_gen2 (developer aid, GPU box only): clock64 deltas between the events
of the LAST step of a run, per compute warp of CTA 0, averaged over the layers.
Events: 0 layer top, 1 layer weights ready, 2 conv MMAs done, 3 z stored / ring stored, 4 barrier 1 passed, 5 residual done,
7 skip MMAs done, 8 barrier 2 passed; per step: 20 step top, 9 layers done, 10 POST1 done,
11 POST2 done, 12 sampled, 13 step end."""
import os, sys
from collections import defaultdict
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lb_wavenet_b200 import _lib, config
from lb_wavenet_b200.engine import GenEngine, TrainEngine
arch = config.load_arch(os.path.join(ROOT, "par", "arch_classic_3x10.json"))
lib = _lib.load()
t = TrainEngine(arch, 1)
t.params.normal_(0, 0.05)
g = GenEngine(arch, 256)
g.load_params(t.params)
g.run(200, seed=0)
torch.cuda.synchronize()
buf = torch.zeros(32 * 2048 + 4 * 1024 + 128, dtype=torch.int64, device="cuda")
lib.wn_debug_trace(buf.data_ptr(), -1)
g.run(200, seed=0)
torch.cuda.synchronize()
lib.wn_debug_trace(None, -1)
b = buf.cpu().numpy()
for warp in (0, 3, 4, 7):
    ev = [(int(x) >> 56, (int(x) >> 40) & 0xffff, int(x) & 0xffffffffff) for x in b[warp * 2048:(warp + 1) * 2048] if x != 0]
    if not ev:
        continue
    d = defaultdict(list)
    for (c0, l0, t0), (c1, l1, t1) in zip(ev, ev[1:]):
        d[(c0, c1)].append(t1 - t0)
    print("warp %d: step %d cycles" % (warp, ev[-1][2] - ev[0][2]))
    for k in sorted(d):
        v = d[k]
        print("   %2d -> %2d : n %3d  avg %7.1f  min %6d  max %6d  total %7d" % (k[0], k[1], len(v), sum(v) / len(v), min(v), max(v), sum(v)))
