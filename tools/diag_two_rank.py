#!/usr/bin/env python
"""Developer aid: the 2-rank NCCL step against the 1-rank step, tensor by tensor (gradients after the all-reduce and weights
after Adam, three steps).  Run once plain and once under torch.distributed.run with 2 ranks; compare the two .npz files.

    python tools/diag_two_rank.py gpurun_out/diag_one.npz
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/diag_two_rank.py gpurun_out/diag_two.npz
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lb_wavenet_b200.data import SlotDealer  # noqa: E402
from lb_wavenet_b200.dist import DistContext  # noqa: E402
from lb_wavenet_b200.tmodel import AdamOptimizer, WaveNetTrain  # noqa: E402
from tests import util  # noqa: E402

arch = dict(util.CLASSIC_SHALLOW, n_lc_in=0, n_lc_out=0, lc_upsample=[], wav_input_type="mu_law_quant")
B, T = 4, 512
ctx = DistContext.from_env("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
torch.cuda.set_device(ctx.local_rank)
net = WaveNetTrain(**arch, batch_sz=B, l2_factor=1e-3, add_summary=False, n_keep_checkpoints=1, ckpt_path="/tmp/nccl.net",
                   resume_step=0, n_valid_total=1, print_interval=0, dist=ctx, init_seed=5, device="cuda:%d" % ctx.local_rank)
net.build()
net.init_vars()
opt = AdamOptimizer(1e-3)
net.bind_optimizer(opt)
eng = net._ensure_engine()
rng = np.random.default_rng(0)
cat = [(int(rng.integers(1, 11)), rng.integers(0, 256, int(rng.integers(300, 900))).astype(np.int32)) for _ in range(9)]
lo, hi = ctx.slot_range(B)
deal = SlotDealer(cat, B, T, net.get_recep_field_sz(), 1, 5, 0, lo, hi, quiet=True)
out = {}
for step in range(3):
    _, w, i = deal.next_batch()
    w, i = torch.as_tensor(w), torch.as_tensor(i)
    net.forward_backward(w, i, True)
    torch.cuda.synchronize()
    for k in eng.reg.params:
        out["g%d_%s" % (step, k)] = eng.view(k, eng.grads).cpu().numpy().copy()
    out["stats%d" % step] = net._gstats.cpu().numpy().copy()
    opt.t += 1
    eng.adam(opt.t, opt.learning_rate, net.l2_factor, n_valid=net._gstats[1:2], beta1=opt.beta1, beta2=opt.beta2, eps=opt.epsilon)
    out["loss%d" % step] = np.float64(net._finish_step(w))
    torch.cuda.synchronize()
    for k in eng.reg.params:
        out["w%d_%s" % (step, k)] = eng.view(k).cpu().numpy().copy()
if ctx.rank == 0:
    np.savez(sys.argv[1], **out)
ctx.barrier()
if ctx.world > 1:
    torch.distributed.destroy_process_group()
