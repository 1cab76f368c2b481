#!/usr/bin/env python
"""Cycles the chain warp of k_gen3 (CTA 0) spends per section of a layer (developer aid, GPU box only).
Sections: 0 publish x (+ wait x_free), 1 wait weight slot, 2 wait x[t-dil] + A fragments, 3 conv (32 mma.sync),
4 gate, 5 publish z (+ wait z_free), 6 residual + x update, 7 slot release, 8 (rest of the layer loop), 9 post-net + sampler."""
import os, sys
os.environ["WN_GEN3"] = "1"  # read once per process by the library
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lb_wavenet_b200 import _lib, config
from lb_wavenet_b200.engine import GenEngine, TrainEngine
arch = config.load_arch(os.path.join(ROOT, "par", "arch_classic_3x10.json"))
lib = _lib.load()
t = TrainEngine(arch, 1)
t.params.normal_(0, 0.05)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = GenEngine(arch, n)
g.load_params(t.params)
g.run(200, seed=0)
torch.cuda.synchronize()
buf = torch.zeros(32 * 2048 + 4 * 1024 + 128, dtype=torch.int64, device="cuda")
lib.wn_debug_trace(buf.data_ptr(), -1)
steps = 500
g.run(steps, seed=0)
torch.cuda.synchronize()
lib.wn_debug_trace(None, -1)
b = buf[:10].cpu().numpy()
L = arch["n_blocks"] * arch["n_block_layers"]
names = ["publish x", "wait weights", "wait old + frags", "conv", "gate", "publish z", "residual", "release", "loop rest", "post-net+sampler"]
for i, nm in enumerate(names):
    per = b[i] / steps / (1 if i >= 8 else L)
    print("%-18s %8.1f cycles per %s" % (nm, per, "step" if i >= 8 else "layer"))
print("per layer total %.1f, per step %.1f" % (sum(b[:8]) / steps / L, sum(b) / steps))
