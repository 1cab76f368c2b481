#!/usr/bin/env python
"""Generator step time at BASELINE configs[3] shapes (256 streams, classic 3x10): developer aid, GPU box only."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lb_wavenet_b200 import config
from lb_wavenet_b200.engine import GenEngine, TrainEngine
arch = config.load_arch(os.path.join(ROOT, "par", "arch_classic_3x10.json"))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
t = TrainEngine(arch, 1)
t.params.normal_(0, 0.05)
g = GenEngine(arch, n)
g.load_params(t.params)
g.run(50, seed=0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.run(steps, seed=0); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("dbg=%s streams %d: %.2f us/step, %.2f M samples/s" % (os.environ.get("WN_GEN_DBG", "0"), n, ms * 1e3 / steps, n * steps / ms / 1e3))
