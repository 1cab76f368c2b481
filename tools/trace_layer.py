#!/usr/bin/env python
"""In-kernel timeline of the persistent layer kernels (developer aid, GPU box only).

    python tools/trace_layer.py [fwd|bwd] [layer]                 (library built with NVCC_EXTRA=-DWN_LAYER_TRACE)
    python tools/trace_layer.py postfwd | postbwd | wgrad     (library built with NVCC_EXTRA=-DWN_POST_TRACE)

Runs BASELINE configs[1] once with tracing off, then one forward (+ backward) with wn_debug_trace pointing at a
device buffer during ONE layer launch, and prints per-role event timelines of CTA 0 (clock64 deltas in cycles).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lb_wavenet_b200 import _lib, config  # noqa: E402
from lb_wavenet_b200.engine import TrainEngine  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "bwd"
layer = int(sys.argv[2]) if len(sys.argv) > 2 else 15
arch = config.load_arch(os.path.join(ROOT, "par", "arch_classic_3x10.json"))
B, T = 32, 16384
eng = TrainEngine(arch, B)
rng = np.random.default_rng(0)
wav = torch.as_tensor(rng.integers(0, 256, (B, T)).astype(np.int32)).cuda()
ids = torch.ones(B, T, dtype=torch.int32).cuda()
lib = _lib.load()
for _ in range(2):
    eng.forward(wav, ids)
    eng.backward()
torch.cuda.synchronize()
buf = torch.zeros(32 * 2048 + 4 * 1024 + 128, dtype=torch.int64, device="cuda")
buf[32 * 2048 + 4096::2] = 2 ** 62
L = len(eng.reg.saves)
if which == "wgrad":
    # layer -3 selects k_wgrad_umma: every launch of the backward logs, the LAST one (SKIP gradient) wins
    eng.forward(wav, ids)
    torch.cuda.synchronize()
    lib.wn_debug_trace(buf.data_ptr(), -3)
    eng.backward_phases(0, 1)
    torch.cuda.synchronize()
    lib.wn_debug_trace(None, -1)
elif which in ("postfwd", "postbwd"):
    # layer -2 selects the post-net kernels (train_umma.cu); both log into the same buffer, so trace one at a time
    if which == "postfwd":
        lib.wn_debug_trace(buf.data_ptr(), -2)
        eng.forward(wav, ids)
        torch.cuda.synchronize()
        lib.wn_debug_trace(None, -1)
    else:
        eng.forward(wav, ids)
        torch.cuda.synchronize()
        lib.wn_debug_trace(buf.data_ptr(), -2)
        eng.backward_phases(0, 1)
        torch.cuda.synchronize()
        lib.wn_debug_trace(None, -1)
elif which == "fwd":
    # tracing stays on for the whole forward: every layer overwrites the buffer, the LAST layer that logged wins;
    # so run the forward with tracing and keep only the wanted layer via phases is not possible -> trace layer L-2
    lib.wn_debug_trace(buf.data_ptr(), layer)
    eng.forward(wav, ids)
    torch.cuda.synchronize()
    lib.wn_debug_trace(None, -1)
else:
    eng.forward(wav, ids)
    torch.cuda.synchronize()
    eng.backward_phases(0, L - layer)      # everything above `layer`
    torch.cuda.synchronize()
    lib.wn_debug_trace(buf.data_ptr(), layer)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.backward_phases(L - layer, L - layer + 1)
    e1.record()
    torch.cuda.synchronize()
    print("traced launch, CUDA events: %.1f us" % (1e3 * e0.elapsed_time(e1)))
    lib.wn_debug_trace(None, -1)
    # the same layer launched three times back to back without tracing
    eng.forward(wav, ids)
    eng.backward_phases(0, L - layer)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    for k in range(3):
        eng.backward_phases(L - layer, L - layer + 1)
        ev[k + 1].record()
    torch.cuda.synchronize()
    print("untraced x3, CUDA events:", ["%.1f us" % (1e3 * ev[k].elapsed_time(ev[k + 1])) for k in range(3)])
    # gaps between back-to-back launches (no events in between): whole backward with every layer traced
    buf2 = torch.zeros_like(buf)
    buf2[32 * 2048 + 4096::2] = 2 ** 62
    eng.forward(wav, ids)
    torch.cuda.synchronize()
    lib.wn_debug_trace(buf2.data_ptr(), -1)
    eng.backward()
    torch.cuda.synchronize()
    lib.wn_debug_trace(None, -1)
    se = buf2[32 * 2048 + 4096:].cpu().numpy().reshape(64, 2)
    se = se[(se[:, 1] > 0) & (se[:, 0] < 2 ** 62)]
    se = se[np.argsort(se[:, 0])]
    print("back-to-back layer launches: body us", np.round((se[:, 1] - se[:, 0]) / 1e3, 1).tolist())
    print("gap (last CTA end -> next kernel first entry) us", np.round((se[1:, 0] - se[:-1, 1]) / 1e3, 1).tolist())
allbuf = buf.cpu().numpy()
ev = allbuf[:32 * 2048].reshape(32, 2048)
cta = allbuf[32 * 2048:32 * 2048 + 4096].reshape(1024, 4)
cta = cta[cta[:, 0] > 0]
if len(cta):
    g0 = cta[:, 0].min()
    print("per-CTA wall clock (ns since first CTA start): n=%d" % len(cta))
    for k in range(0, len(cta), max(1, len(cta) // 40)):
        print("  cta %4d sm %3d start %7d end %7d" % (k, cta[k, 2], cta[k, 0] - g0, cta[k, 1] - g0))
    print("  last end %d ns" % (cta[:, 1].max() - g0))
    if (cta[:, 3] > 0).all():
        print("  kernel entry -> init done: min %d max %d ns; first entry -> first start %d ns; entry spread %d ns" % (
            (cta[:, 0] - cta[:, 3]).min(), (cta[:, 0] - cta[:, 3]).max(), g0 - cta[:, 3].min(), cta[:, 3].max() - cta[:, 3].min()))
names = {1: "prod:wait_free", 2: "prod:got_free", 3: "mma:issueA", 4: "mma:issueB", 5: "e1:begin", 14: "e1:in_full",
         6: "e1:v_full", 15: "e1:math_done", 7: "e1:pre_bar", 8: "e1:post_bar", 9: "e2:begin", 10: "e2:acc_full",
         11: "e2:pre_bar", 12: "e2:post_bar", 13: "e2:stored",
         16: "mma:doneA", 17: "mma:doneB", 18: "e0:begin", 19: "e0:in_full", 20: "e0:done",
         30: "k:entry", 31: "k:init_done", 32: "k:role_done", 33: "k:all_done", 34: "k:bar_init", 35: "k:bias",
         36: "k:tmem_alloc", 37: "k:synced"}
if which == "wgrad":
    names = {1: "prod:wait_empty", 2: "prod:got_empty", 3: "mma:wait_full", 16: "mma:stage_full", 5: "epi:wait_acc",
             6: "epi:acc_full", 30: "k:entry", 31: "k:pdl_wait_done", 32: "k:role_done", 33: "k:all_done"}
if which in ("postfwd", "postbwd"):
    names = {1: "prod:wait_empty", 2: "prod:got_empty", 3: "mma:wait_acc_empty", 4: "mma:gemm_begin", 16: "mma:stage_full",
             17: "mma:gemm_issued", 5: "epi:wait_acc", 6: "epi:acc_full", 7: "epi:acc_released", 8: "epi:staging_free",
             9: "epi:staged"}
rows = []
for w in range(32):
    for x in ev[w]:
        if x == 0:
            continue
        rows.append((int(x) & 0xffffffffff, w, (int(x) >> 56) & 0xff, (int(x) >> 40) & 0xffff))
rows.sort()
if not rows:
    print("no events")
    sys.exit(0)
t0 = rows[0][0]
maxtile = int(os.environ.get("TRACE_TILES", "12"))
lo = int(os.environ.get("TRACE_FROM", "6"))
for t, w, c, tile in rows:
    if lo <= tile < lo + maxtile or c >= 30:
        print("%9d  w%-2d tile %3d  %s" % (t - t0, w, tile, names.get(c, str(c))))
# per-event mean period
last = {}
per = {}
for t, w, c, tile in rows:
    k = (w, c)
    if k in last:
        per.setdefault(k, []).append(t - last[k])
    last[k] = t
print("mean period per (warp, event):")
for k in sorted(per):
    print("  w%-2d %-16s %8.0f cycles over %d" % (k[0], names.get(k[1], str(k[1])), np.mean(per[k]), len(per[k])))
print("total span %d cycles, tiles %d" % (rows[-1][0] - t0, max(r[3] for r in rows) + 1))
