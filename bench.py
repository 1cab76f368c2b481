#!/usr/bin/env python
"""Benchmark of the lb-wavenet hot path on B200 (contract: see the task's bench.py section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric: training output timesteps/s (= B*(T-1) loss positions per optimiser step / step time,
forward + backward + gradient all-reduce + Adam), BASELINE.json configs[1]: classic 3x10 stack,
R=D=32, S=P=256, 32 slots per GPU, slice_sz 16384, bf16 operands.  N>1 (torchrun): the B=32*N
slots are sharded 32 per GPU (weak scaling, BASELINE.json configs[2] at N=8).
Extra keys: `gen` = BASELINE configs[3], batched incremental generation audio samples/s on 1 GPU (256 streams x 160 000
steps, its own `roofline` and `e2e`); `configs0` = BASELINE configs[0] (reference par/arch1.json + par/par1.json: one
10 x 512 training step and 16 000 generation steps x 10 streams); `e2e.loader` = H2D bytes per timestep, achieved copy
GB/s and the dealer's host time; `roofline.hbm` / `roofline.tensor` = BOTH fractions of the dominant kernel.
`--impl reference` times the CPU port of the reference graph on a bounded sample and says so in `config`.

Timing: W warm-up steps; 3 untimed steps with CUDA events around EVERY kernel launch (`kernel_shares`); then exactly K
timed steps bracketed by barrier + synchronize, with live CUDA events only around the launches of the dominant kernel
(`roofline.achieved`; an event pair costs ~1 us of stream time per launch, ~0.15 ms per step if every launch is timed);
then K end-to-end steps through WaveNetTrain.train_step with pinned host inputs and the loss read back, and K through the
reference's own loop (MaskedSliceWav -> net.run), which is what `e2e` reports.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ARCH_FILE = os.path.join(ROOT, "par", "arch_classic_3x10.json")
SLOTS_PER_GPU = 32
SLICE_SZ = 16384
GEN_STREAMS = 256
GEN_STEPS = 160000   # BASELINE configs[3]: 10 s at 16 kHz per stream


def train_flop_per_timestep(a) -> float:
    """fwd = L*(8RD + 2DR + 2DS [+4GD]) + 2SP + 2PQ ; train = 3 x fwd."""
    L = a["n_blocks"] * a["n_block_layers"]
    R, D, S, P, Q, G = a["n_res"], a["n_dil"], a["n_skip"], a["n_post"], a["n_quant"], a["n_gc_embed"]
    fwd = L * (8 * R * D + 2 * D * R + 2 * D * S + 4 * G * D) + 2 * S * P + 2 * P * Q
    return 3.0 * fwd


def post_fwd_flop_per_timestep(a) -> float:
    """skip GEMM (concat-K over layers) + POST1 + POST2: the dominant forward kernel."""
    L = a["n_blocks"] * a["n_block_layers"]
    D, S, P, Q = a["n_dil"], a["n_skip"], a["n_post"], a["n_quant"]
    return 2.0 * L * D * S + 2.0 * S * P + 2.0 * P * Q


def synth_slots(n_slots, T, n_batches, seed, recep_field):
    """Synthetic 16 kHz audio per slot: 3 sinusoids + noise, mu-law encoded by the loader's own
    dealer from in-memory 'files' of length U[2F, 8F], ids from the real slot dealer."""
    from lb_wavenet_b200.data import SlotDealer
    files = synth_files(max(8, n_slots), seed, recep_field)
    d = SlotDealer(files, n_slots, T, recep_field, 1, seed, 0, quiet=True)
    return [d.next_batch()[1:] for _ in range(n_batches)]


def synth_files(n_files, seed, recep_field):
    """In-memory synthetic 'files' (voice id, mu-law codes): lengths U[2F, 8F], 3 sinusoids + noise."""
    rng = np.random.default_rng(seed)
    files = []
    for i in range(n_files):
        n = int(rng.integers(2 * recep_field, 8 * recep_field))
        t = np.arange(n)
        f = rng.uniform(100, 4000, 3) / 16000.0
        x = sum(0.3 * np.sin(2 * np.pi * fi * t + rng.uniform(0, 6.28)) for fi in f) + rng.normal(0, 0.05, n)
        x = np.clip(x, -1, 1)
        q = (np.sign(x) * np.log1p(255 * np.abs(x)) / np.log1p(255) + 1) * 0.5 * 255 + 0.5  # mu-law codes
        files.append((int(rng.integers(1, 100)), q.astype(np.int32)))
    return files


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(arch, steps, warmup, n_threads=None):
    """The reference's CPU implementation of the path: the oracle port (torch CPU fp32, op for op
    as tmodel.py:292-340 + TF Adam), timed on a bounded sample of the workload with all host threads."""
    import torch
    from oracle import wavenet_oracle as O
    n_threads = n_threads or os.cpu_count() or 1
    torch.set_num_threads(n_threads)
    a = O.Arch(arch["n_blocks"], arch["n_block_layers"], arch["n_quant"], arch["n_res"], arch["n_dil"],
               arch["n_skip"], arch["n_post"], arch["n_gc_embed"], arch["n_gc_category"], bool(arch["use_bias"]))
    B, T = 8, 8192
    p = O.init_params(a, B, seed=0)
    wav, ids = synth_slots(B, T, 1, 1234, a.recep_field())[0]
    ids = np.maximum(ids, 1)  # bounded sample: keep every position valid so the work is the same per step
    m = {k: np.zeros_like(v) for k, v in p.items() if v.dtype.kind == "f" and not k.startswith("SAVE")}
    v2 = {k: np.zeros_like(v) for k, v in m.items()}
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        grads, L, fwd = O.train_step_autograd(a, p, wav, ids, 1e-3, torch.float32)
        for k in m:
            g = grads[k].astype(np.float32)
            p[k], m[k], v2[k] = O.adam_tf_step(p[k], g, m[k], v2[k], it + 1, 1e-3)
            p[k] = p[k].astype(np.float32)
        for li, name in enumerate(n for n in p if n.startswith("SAVE")):
            p[name] = fwd.new_save[li].numpy().astype(np.float32)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    per_step = float(np.mean(times))
    return dict(value=B * (T - 1) / per_step, ms_per_step=per_step * 1e3, cores=n_threads,
                sample="%d slots x %d timesteps per step, %d steps, torch-CPU fp32 port of the reference graph "
                       "(TensorFlow 1.x not installable)" % (B, T, steps), B=B, T=T)


def gen_leg(arch, params, n_streams, n_steps, dev, lib, workload, gc_ids=None, e2e=True):
    """Batched incremental generation: `value` = streams x steps / device time of ONE persistent launch (CUDA events);
    `e2e` = the same through the host call including the device -> host copy of the sampled codes (wall clock around
    run + .cpu()).  Roofline per SURVEY 8(d): ring-buffer bytes per stream-step = L * 2 * R * 2 (read x[t-dil], write
    x[t], bf16) against the measured HBM copy peak -- the path is latency-bound (L dependent layer hops per sample), so
    us_per_step is the figure that matters."""
    import torch
    from lb_wavenet_b200.engine import GenEngine
    g = GenEngine(arch, n_streams, dev)
    g.load_params(params, gc_ids)
    g.run(50, seed=0)
    torch.cuda.synchronize()
    lib.wn_launch_count_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    codes = g.run(n_steps, seed=0)
    e1.record()
    torch.cuda.synchronize()
    gms = e0.elapsed_time(e1)
    launches = int(lib.wn_launch_count_reset())
    sps = n_streams * n_steps / (gms * 1e-3)
    L = arch["n_blocks"] * arch["n_block_layers"]
    ring_bytes = L * 2 * arch["n_res"] * 2
    hbm_peak = 6650.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        with open(pk) as f:
            hbm_peak = json.load(f)["hbm_gbs"]
    out = {"metric": "batched gen audio samples/s", "value": sps, "unit": "samples/s", "streams": n_streams,
           "steps": n_steps, "us_per_step": gms * 1e3 / n_steps, "realtime_multiple_per_stream": sps / n_streams / 16000.0,
           "gpu_launches": launches, "workload": workload,
           "roofline": {"bound": "hbm", "achieved": sps * ring_bytes / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": sps * ring_bytes / 1e9 / hbm_peak, "traffic": None,
                        "algorithmic_bytes_per_stream_step": ring_bytes,
                        "note": "latency-bound chain of L dependent layer steps per sample, not a bandwidth limit"}}
    if e2e:
        del codes
        g.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        host_codes = g.run(n_steps, seed=0).cpu()
        dt = time.perf_counter() - t0
        out["e2e"] = {"value": n_streams * n_steps / dt, "unit": "samples/s", "h2d_bytes_per_step": 0,
                      "d2h_bytes_per_step": int(host_codes.numel() * host_codes.element_size()),
                      "note": "one step = the whole %d-step run: GenEngine.run + device -> host copy of the int32 codes, "
                              "host wall clock" % n_steps}
    del g
    return out


def configs0_leg(dev, lib):
    """BASELINE.json configs[0]: par/arch1.json + par/par1.json of the reference (5x10 stack, R = D = 32, S = P = 512,
    global conditioning 17 x 377; batch 10, slice 512, l2 = lr = 1e-3) -- one stage-wise training step and 1 s
    (16 000 steps) of incremental generation for the 10 streams.  A launch-latency-sized workload (5 120 timesteps per
    step): reported beside the headline, not as it."""
    import torch
    from lb_wavenet_b200 import config
    from lb_wavenet_b200.tmodel import AdamOptimizer, WaveNetTrain
    arch = config.load_arch(os.path.join(ROOT, "par", "arch_c1_5x10_gc.json"))
    par = config.load_par(os.path.join(ROOT, "par", "par_c1.json"))
    B, T = par["batch_sz"], par["slice_sz"]
    net = WaveNetTrain(**arch, batch_sz=B, l2_factor=par["l2_factor"], add_summary=False, n_keep_checkpoints=1,
                       ckpt_path="/tmp/bench0.net", resume_step=0, n_valid_total=1, print_interval=0, init_seed=0,
                       device=dev)
    net.build()
    net.init_vars()
    opt = AdamOptimizer(par["learning_rate"])
    F = net.get_recep_field_sz()
    files = synth_files(16, 777, F)
    files = [(1 + (v % arch["n_gc_category"]), q) for v, q in files]
    from lb_wavenet_b200.data import SlotDealer
    d = SlotDealer(files, B, T, F, 1, 5, 0, quiet=True)
    batches = [d.next_batch()[1:] for _ in range(4)]
    pinned = [(torch.as_tensor(w).pin_memory(), torch.as_tensor(i).pin_memory()) for w, i in batches]
    devb = [(w.to(dev), i.to(dev)) for w, i in pinned]
    K = 50
    for s_ in range(5):
        net.train_step(*devb[s_ % 4], opt, want_loss=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s_ in range(K):
        net.train_step(*devb[s_ % 4], opt, want_loss=False)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    t0 = time.perf_counter()
    for s_ in range(K):
        loss = net.train_step(*pinned[s_ % 4], opt, want_loss=True)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / K
    out = {"workload": "BASELINE.json configs[0]: reference par/arch1.json (normalised: par/arch_c1_5x10_gc.json) + "
                       "par/par1.json, batch 10 x slice 512, one stage-wise training step; 16 000 generation steps x 10 streams",
           "train": {"value": B * (T - 1) / (ms * 1e-3), "unit": "timesteps/s", "ms_per_step": ms,
                     "e2e": {"value": B * (T - 1) / (e2e_ms * 1e-3), "ms_per_step": e2e_ms, "loss": loss,
                             "h2d_bytes_per_step": 8 * B * T, "d2h_bytes_per_step": 32}}}
    gc_ids = np.arange(1, 11, dtype=np.int32)
    out["gen"] = gen_leg(net._arch_dict, net.engine.params, 10, 16000, dev, lib, "10 streams x 16 000 steps (1 s)", gc_ids)
    return out


_JSON_OUT = None


def _emit(text):
    out = _JSON_OUT if _JSON_OUT is not None else sys.stdout
    out.write(text + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-gen", action="store_true", help="skip the generation leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-configs0", action="store_true", help="skip the BASELINE configs[0] (arch1 + par1) leg")
    ap.add_argument("--gen-steps", type=int, default=GEN_STEPS, help="generation steps per stream (configs[3]: 160000 = 10 s)")
    ap.add_argument("--slots", type=int, default=None)
    ap.add_argument("--slice", type=int, default=None)
    ap.add_argument("--workload", default="classic", choices=["classic", "wide"],
                    help="classic = BASELINE.json configs[1] (the metric's configuration); wide = configs[4] "
                         "(4x10, R=D=128, S=P=512, 8 slots/GPU x 32768): a second, non-headline line")
    args = ap.parse_args()
    global ARCH_FILE
    wide = args.workload == "wide"
    if wide:
        ARCH_FILE = os.path.join(ROOT, "par", "arch_wide_4x10.json")
    args.slots = args.slots or (8 if wide else SLOTS_PER_GPU)
    args.slice = args.slice or (32768 if wide else SLICE_SZ)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    from lb_wavenet_b200 import config
    arch = config.load_arch(ARCH_FILE)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries ONE JSON line: keep a handle on the real stdout for it and point file descriptor 1 at stderr, so that
    # whatever a library prints on the way (NCCL's version banner under NCCL_DEBUG=VERSION is a plain printf) lands there
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    base = {
        "metric": "train output timesteps/s", "unit": "timesteps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": ("BASELINE.json configs[4]: wide 4x10 (R=D=128,S=P=512)" if wide else
                                "BASELINE.json configs[1]: classic 3x10 (R=D=32,S=P=256)") +
                               ", %d slots/GPU x slice_sz %d, one stage-wise training step (fwd+bwd+allreduce+Adam), "
                               "consecutive stages with carried SAVE state" % (args.slots, args.slice),
                   "arch_file": "par/" + os.path.basename(ARCH_FILE), "slots_per_gpu": args.slots, "slice_sz": args.slice,
                   "global_slots": args.slots * args.gpus, "parallelism": "dp%d over slots" % args.gpus,
                   "l2_flush": "inputs and activation stash (>4 GB/step) far exceed the 126 MB L2"},
    }

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = cpu_reference_run(arch, args.steps, args.warmup)
        line = dict(base)
        line["config"] = dict(base["config"])
        line["config"]["workload"] += (" -- REFERENCE ARM: CPU port of the reference graph (torch fp32; TensorFlow 1.x is "
                                       "not installable), timed on a BOUNDED SAMPLE of this workload: %d slots x %d "
                                       "timesteps per step, every position valid (throughput is ~linear in slots x "
                                       "timesteps)" % (r["B"], r["T"]))
        line["config"]["reference_sample"] = {"slots": r["B"], "slice_sz": r["T"], "dtype": "f32",
                                              "same_shape_as_gpu_arm": False}
        line.update({"impl": "reference", "value": r["value"], "ms_per_step": r["ms_per_step"],
                     "dtype": "f32", "gpu_launches": 0,
                     "cpu_baseline": {"value": r["value"], "unit": "timesteps/s", "cores": r["cores"], "kind": "port",
                                      "sample": r["sample"]},
                     "e2e": {"value": r["value"], "unit": "timesteps/s", "h2d_bytes_per_step": 0,
                             "d2h_bytes_per_step": 0}})
        _emit(json.dumps(line))
        return 0

    import torch
    from lb_wavenet_b200 import _lib
    from lb_wavenet_b200.dist import DistContext
    from lb_wavenet_b200.tmodel import AdamOptimizer, WaveNetTrain
    lib = _lib.load()
    ctx = DistContext.from_env("nccl" if world > 1 else None)
    if world == 1:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", ctx.local_rank)
    B_total, T = args.slots * max(world, 1), args.slice
    net = WaveNetTrain(**arch, batch_sz=B_total, l2_factor=1e-3, add_summary=False, n_keep_checkpoints=1,
                       ckpt_path="/tmp/bench.net", resume_step=0, n_valid_total=1, print_interval=0, dist=ctx,
                       init_seed=0, device=str(dev))
    # the reference-facing loop (train.py:133-186,216-240): dataset -> get_op() -> net.build(*ops) -> sess.run stand-in
    from lb_wavenet_b200 import data as wdata
    F = net.get_recep_field_sz()
    dset = wdata.MaskedSliceWav(None, None, 16000, T, 2, 0, 1, B_total, 1, "/tmp/bench.dset", 0,
                                dist=ctx if world > 1 else None, device=str(dev), random_seed=99)
    dset.init_sample_catalog(entries=synth_files(max(8, B_total), 4321, F))  # same catalog on every rank
    dset.set_receptive_field_size(F)
    dset.build()
    _, *data_ops = dset.get_op()
    gv_op, loss_op = net.build(*data_ops)
    net.init_vars()
    opt = AdamOptimizer(1e-3)
    apply_op = opt.apply_gradients(gv_op)
    n_batches = 2
    host_batches = synth_slots(args.slots, T, n_batches, 1234 + rank, F)
    pinned = [(torch.as_tensor(w).pin_memory(), torch.as_tensor(i).pin_memory()) for w, i in host_batches]
    dev_batches = [(w.to(dev), i.to(dev)) for w, i in pinned]
    torch.cuda.synchronize()

    def sync_all():
        torch.cuda.synchronize()
        ctx.barrier()

    # ---- device-resident timing: `value` ---------------------------------------------------
    import ctypes as C
    cats = ["prep_embed_save", "layer_fwd", "post_fwd_loss", "post_bwd", "layer_bwd", "layer_bwd_data",
            "wgrad", "pre_gc_bwd", "adam", "gen"]
    for s in range(args.warmup):
        w, i = dev_batches[s % n_batches]
        net.train_step(w, i, opt, want_loss=False)
    sync_all()
    # kernel_shares: per-category CUDA-event timing of every launch over 3 extra UNTIMED steps (an event pair costs
    # ~1 us of stream time per launch, ~0.15 ms per step with ~90 launches: too much to leave inside `value`)
    n_prof = 3
    lib.wn_prof_enable(1)
    for s in range(n_prof):
        w, i = dev_batches[s % n_batches]
        net.train_step(w, i, opt, want_loss=False)
    prof_ms = (C.c_double * 16)()
    prof_n = (C.c_int64 * 16)()
    lib.wn_prof_collect(prof_ms, prof_n)
    lib.wn_prof_enable(0)
    shares = {c: {"ms_per_step": prof_ms[k] / n_prof, "launches_per_step": prof_n[k] / n_prof}
              for k, c in enumerate(cats) if prof_n[k] > 0}
    dom = max(shares, key=lambda c: shares[c]["ms_per_step"]) if shares else None
    sync_all()
    clocks = ClockSampler(ctx.local_rank)
    if rank == 0:
        clocks.start()
    lib.wn_launch_count_reset()
    # the timed region keeps live CUDA events around the launches of the DOMINANT kernel only (roofline.achieved)
    if dom is not None:
        lib.wn_prof_enable(1 << (cats.index(dom) + 1))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        w, i = dev_batches[s % n_batches]
        net.train_step(w, i, opt, want_loss=False)
    e1.record()
    sync_all()
    ms_total = e0.elapsed_time(e1)
    launches = int(lib.wn_launch_count_reset())
    dom_ms = (C.c_double * 16)()
    dom_n = (C.c_int64 * 16)()
    lib.wn_prof_collect(dom_ms, dom_n)
    lib.wn_prof_enable(0)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = B_total * (T - 1) / (ms_per_step * 1e-3)

    # ---- end-to-end through the public API with host buffers: `e2e` --------------------------------
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = None
    for s in range(args.steps):
        w, i = pinned[s % n_batches]
        loss = net.train_step(w, i, opt, want_loss=True)  # H2D of the inputs + D2H of the loss inside
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    direct_ms = float(t.item()) / args.steps
    # (b) the reference's own loop: net.run([apply_grads_op, loss_op]) pulling from MaskedSliceWav -- the loader thread
    # deals windows into pinned buffers and copies them on its own stream (double buffered), so the H2D copy of step
    # i+1 overlaps step i; the loss is still read back (and the host blocked) every step
    dset.init_vars()  # start the loader thread
    for s in range(3):
        net.run([apply_op, loss_op])
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        _, loss = net.run([apply_op, loss_op])
    e1.record()
    sync_all()
    loader = dset.loader_stats()
    dset._shutdown()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item()) / args.steps
    clk = clocks.stop() if rank == 0 else None  # sampled over all timed regions
    e2e = {"value": B_total * (T - 1) / (e2e_ms * 1e-3), "unit": "timesteps/s", "ms_per_step": e2e_ms,
           # uint8 mu-law code + int32 id per timestep, counted from the pinned tensors the loader copies
           "h2d_bytes_per_step": int(loader.get("h2d_bytes_per_batch", 5 * args.slots * T) * max(world, 1)),
           "d2h_bytes_per_step": int(8 * _lib.WN_NSTATS * max(world, 1)), "loss": loss,
           "api": "MaskedSliceWav.get_op() -> WaveNetTrain.build(*ops) -> net.run([apply_grads_op, loss_op]) (reference "
                  "train.py:182-240); inputs dealt on the host into pinned buffers, H2D on the loader's copy stream",
           "direct_train_step_ms": direct_ms,
           "loader": dict(loader, note="h2d_gbs: CUDA events on the loader's copy stream around the two H2D copies of a "
                                       "batch (pinned -> device); deal_ms_per_batch: host time of the C slot dealer "
                                       "(all global slots replayed, local slots materialised)"),
           "direct_note": "WaveNetTrain.train_step(pinned host tensors): the same step with the H2D copy serialised on "
                          "the compute stream"}

    if rank != 0:
        ctx.barrier()
        import torch.distributed as dist
        dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (its launches were timed live, inside the timed region) ----
    if dom is not None and dom_n[cats.index(dom)] > 0:
        k = cats.index(dom)
        shares[dom] = {"ms_per_step": dom_ms[k] / args.steps, "launches_per_step": dom_n[k] / args.steps,
                       "timed_in": "timed region (the other categories: %d untimed profiling steps)" % n_prof}
    peaks = {"bf16_sustained": 1590.0, "bf16_burst": 1590.0, "hbm_gbs": 6650.0,
             "source": "fallback figures of B200_PROFILING.md (MEASURED_PEAKS.json absent)"}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        with open(pk) as f:
            mp = json.load(f)
        peaks = {"bf16_sustained": mp.get("bf16_tflops_sustained", mp["bf16_tflops"]), "bf16_burst": mp["bf16_tflops"],
                 "hbm_gbs": mp["hbm_gbs"], "source": "MEASURED_PEAKS.json"}
    rows = args.slots * T
    L_, R_, D_ = arch["n_blocks"] * arch["n_block_layers"], arch["n_res"], arch["n_dil"]
    flop_by_cat = {
        "post_fwd_loss": post_fwd_flop_per_timestep(arch) * rows,
        "post_bwd": post_fwd_flop_per_timestep(arch) * rows,       # dgrad chain: same contractions transposed
        "wgrad": train_flop_per_timestep(arch) / 3.0 * rows,          # every contraction once more
        "layer_fwd": L_ * (8 * R_ * D_ + 2 * D_ * R_) * rows,         # conv taps + residual (SURVEY 8d formula)
        # fused R = 32 backward: data + weight gradients of conv and residual in one kernel; wide path: the recomputed
        # conv + dz = dx.RESIDUAL^T here, the data gradient under layer_bwd_data, weight gradients under wgrad
        "layer_bwd": L_ * (8 * R_ * D_ + 2 * D_ * R_) * rows * (2 if R_ < 64 else 1),
        "layer_bwd_data": L_ * (8 * R_ * D_) * rows,
    }
    # ALGORITHMIC HBM bytes per timestep (bf16), independent of how the kernel stores its intermediates:
    #   layer_fwd : read x_l (R), write z_l (D) and x_{l+1} (R)                      = 192 B per layer at R = D = 32
    #   layer_bwd : read x_l (R), dz_l (D), dx_{l+1} (R), write dx_l (R)             = 256 B per layer
    # (the fused backward also re-reads x_l shifted by dil, an L2 hit: 320 B, reported as implementation_bytes_per_launch)
    hbm_by_cat = {"layer_fwd": L_ * (2 * R_ + D_) * 2, "layer_bwd": L_ * (3 * R_ + D_) * 2,
                  "layer_bwd_data": L_ * (4 * D_ + 2 * R_) * 2}
    impl_by_cat = {"layer_bwd": L_ * (4 * R_ + D_) * 2} if R_ < 64 else {}
    tensor_bound = ("post_fwd_loss", "post_bwd", "wgrad") + (("layer_fwd", "layer_bwd", "layer_bwd_data") if R_ >= 64 else ())
    roofline = None
    if dom is not None:
        dms = shares[dom]["ms_per_step"]
        n_launch = max(1.0, shares[dom]["launches_per_step"])
        tf_ach = flop_by_cat[dom] / (dms * 1e-3) / 1e12 if dom in flop_by_cat else None
        gb_ach = hbm_by_cat[dom] * rows / (dms * 1e-3) / 1e9 if dom in hbm_by_cat else None
        traffic = None  # measured DRAM bytes per launch of this kernel (one ncu --set full capture, profiles/)
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(dom, {}).get("dram_bytes_per_launch")
        if dom in tensor_bound or gb_ach is None:
            roofline = {"kernel": dom, "bound": "tensor", "achieved": tf_ach, "peak": peaks["bf16_sustained"],
                        "unit": "TFLOP/s", "frac": tf_ach / peaks["bf16_sustained"], "traffic": traffic,
                        "peak_source": peaks["source"] + " bf16_tflops_sustained (kernel timed inside a long step)"}
        else:
            roofline = {"kernel": dom, "bound": "hbm", "achieved": gb_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": gb_ach / peaks["hbm_gbs"], "traffic": traffic,
                        "algorithmic_bytes_per_launch": hbm_by_cat[dom] * rows / n_launch,
                        "peak_source": peaks["source"] + " hbm_gbs (measured copy bandwidth)"}
            if dom in impl_by_cat:
                roofline["implementation_bytes_per_launch"] = impl_by_cat[dom] * rows / n_launch
        roofline.update({"us_per_launch": dms * 1e3 / n_launch, "ms_per_step": dms})
        # both fractions of the dominant kernel, whichever bound is reported above (tensor: against the BURST peak,
        # BASELINE.md section 3)
        if tf_ach is not None:
            roofline["tensor"] = {"achieved_tflops": tf_ach, "peak_burst": peaks["bf16_burst"],
                                  "frac_of_burst": tf_ach / peaks["bf16_burst"]}
        if gb_ach is not None:
            roofline["hbm"] = {"achieved_gbs": gb_ach, "peak": peaks["hbm_gbs"], "frac": gb_ach / peaks["hbm_gbs"]}
    whole = value * train_flop_per_timestep(arch) / 1e12

    line = dict(base)
    line.update({"value": value, "ms_per_step": ms_per_step, "e2e": e2e, "gpu_launches": launches,
                 "clocks": clk, "roofline": roofline,
                 "whole_step_tflops": whole,
                 "whole_step_frac_of_bf16_peak": whole / peaks["bf16_burst"] / max(world, 1),   # burst peak, BASELINE.md 3
                 "whole_step_frac_of_bf16_sustained": whole / peaks["bf16_sustained"] / max(world, 1),
                 "kernel_shares": shares})

    # ---- generation legs (1 GPU, rank 0): batched incremental generation samples/s ---------------------
    if not args.no_gen and world == 1:
        try:
            # BASELINE configs[3]: 256 independent streams x 160 000 steps (10 s each) from the 3x10 stack
            line["gen"] = gen_leg(net._arch_dict, net.engine.params, GEN_STREAMS, args.gen_steps, str(dev), lib,
                                  "BASELINE.json configs[3]: %d streams x %d steps, 3x10 stack, ring buffers in HBM"
                                  % (GEN_STREAMS, args.gen_steps))
            # throughput grows with the batch until all 148 SMs hold a stream group (one CTA per 16 streams)
            g4 = gen_leg(net._arch_dict, net.engine.params, 4 * GEN_STREAMS, 2000, str(dev), lib, "", e2e=False)
            line["gen"]["at_%d_streams" % (4 * GEN_STREAMS)] = g4["value"]
        except Exception as e:  # the training line must survive a generator failure
            line["gen"] = {"error": repr(e)}
    del net, dset
    torch.cuda.empty_cache()
    if not args.no_configs0 and world == 1 and not wide:
        try:
            line["configs0"] = configs0_leg(str(dev), lib)
        except Exception as e:
            line["configs0"] = {"error": repr(e)}

    # ---- CPU baseline beside it (bounded sample) ------------------------------------------------------
    if not args.no_cpu and world == 1:
        r = cpu_reference_run(arch, 3, 1)
        line["cpu_baseline"] = {"value": r["value"], "unit": "timesteps/s", "cores": r["cores"], "kind": "port",
                                "sample": r["sample"]}
    _emit(json.dumps(line))
    if world > 1:
        ctx.barrier()
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
