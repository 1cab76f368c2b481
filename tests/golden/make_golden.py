#!/usr/bin/env python
"""Regenerates the fixtures under tests/golden/.

  readme_influence.json  -- the known-answer diagram of the reference README (README.md:70-85,
      images/wavenet_influence.png): dilations [1,2,4,8]*3, every filter [0.5, 0.5], input
      ...0,0 | 4096,4096...; the values below are the ones legible in the diagram,
      typed in by hand -- NOT produced by the oracle.
  oracle_tiny_forward.npz -- outputs of the oracle itself on a seeded tiny configuration; pins the
      oracle (and through it the CUDA path) against silent regressions.  The reference cannot be
      imported here (TensorFlow 1.x), so this is an oracle-generated vector, labelled as such.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import wavenet_oracle as O  # noqa: E402


def main():
    readme = {
        "dilations": [1, 2, 4, 8] * 3,
        "filter": [0.5, 0.5],
        "input_left": 0, "input_right": 4096,
        # top row, first values right of the junction (t = 0 is the first 4096-valued input)
        "top_row_from_junction": [1, 4, 10, 20, 35, 56, 84, 120, 165, 220, 286, 364, 455, 560, 680, 816, 966,
                                  1128, 1300, 1480, 1666, 1856, 2048, 2240, 2430, 2616],
        "top_row_tail": [4095, 4096],
        "layer1_around_junction": [0, 2048, 4096],
        "layer2_around_junction": [0, 1024, 2048, 3072, 4096],
        "steps_to_saturate": 45,
    }
    with open(os.path.join(HERE, "readme_influence.json"), "w") as f:
        json.dump(readme, f, indent=1)

    a = O.Arch(2, 3, 256, 16, 16, 32, 32, n_gc_embed=5, n_gc_category=7)
    B, T = 2, 40
    p = O.init_params(a, B, seed=123, bias_scale=0.3)
    rng = np.random.default_rng(5)
    wav = rng.integers(0, 256, (B, T))
    ids = rng.integers(0, 8, (B, T))
    grads, L, fwd = O.train_step_autograd(a, p, wav, ids, 1e-3, torch.float64)
    np.savez_compressed(os.path.join(HERE, "oracle_tiny_forward.npz"), wav=wav, ids=ids,
                        logits=fwd.logits.detach().numpy(), total=float(L.total), xent_sum=float(L.xent_sum),
                        n_valid=L.n_valid, diff_sum=L.diff_sum, l2=float(L.l2),
                        grad_PRE=grads["PRE"], grad_SIGNAL_1_2=grads["SIGNAL_1_2"], grad_POST2=grads["POST2"],
                        save_last=fwd.new_save[-1].numpy())


if __name__ == "__main__":
    main()
