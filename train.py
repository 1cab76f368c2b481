#!/usr/bin/env python
"""Training entry point: same command line as the reference's train.py (train.py:12-65).

    train.py [-tf FILE] [-pd DIR] [-rs INT] [-s] [-cpu] [-tb DIR] [-si INT] [-pi INT] [-tdb] [-te]
             [-ms INT] [-bs INT] [-ss INT] [-l2 F] [-lr F] [-gc INT]
             CKPT_PATH_PFX ARCH_FILE PAR_FILE SAMPLES_FILE

Launch one process per GPU with torchrun for data-parallel training over the B slots.
Flags that only made sense for TensorFlow (-cpu, -te, -tdb, -pd, -tb, -s) are accepted and inert:
there is no CPU path by design.  --timeline-file writes per-kernel-category CUDA-event timings of
step 5 as Chrome-trace JSON (the reference dumps a TF timeline at the same step, train.py:225-238,
but also takes a second optimiser step there -- that quirk is not reproduced).
"""
import argparse
import json
from sys import stderr


def get_args(argv=None):
    p = argparse.ArgumentParser(description="WaveNet")
    p.add_argument("--timeline-file", "-tf", type=str, help="Enable profiling and write info to <timeline_file>")
    p.add_argument("--prof-dir", "-pd", type=str, metavar="DIR", help="(accepted, unused: TensorFlow profiler)")
    p.add_argument("--resume-step", "-rs", type=int, metavar="INT",
                   help="Resume training from CKPT_DIR/<ckpt_pfx>-<resume_step>.{meta,index,data-..}")
    p.add_argument("--add-summary", "-s", action="store_true", default=False, help="(accepted, unused)")
    p.add_argument("--cpu-only", "-cpu", action="store_true", default=False, help="(accepted, unused: no CPU path)")
    p.add_argument("--tb-dir", "-tb", type=str, metavar="DIR", help="(accepted, unused)")
    p.add_argument("--save-interval", "-si", type=int, default=1000, metavar="INT",
                   help="Save a checkpoint after this many steps each time")
    p.add_argument("--progress-interval", "-pi", type=int, default=10, metavar="INT",
                   help="Print a progress message at this interval")
    p.add_argument("--tf-debug", "-tdb", action="store_true", default=False, help="(accepted, unused)")
    p.add_argument("--tf-eager", "-te", action="store_true", default=False, help="(accepted, unused)")
    p.add_argument("--max-steps", "-ms", type=int, default=1e20, help="Maximum number of training steps")
    # training parameter overrides
    p.add_argument("--batch-size", "-bs", type=int, metavar="INT", help="Batch size (overrides PAR_FILE setting)")
    p.add_argument("--slice-size", "-ss", type=int, metavar="INT", help="Slice size (overrides PAR_FILE setting)")
    p.add_argument("--l2-factor", "-l2", type=float, metavar="FLOAT", help="Loss = Xent loss + l2_factor * l2_loss")
    p.add_argument("--learning-rate", "-lr", type=float, metavar="FLOAT",
                   help="Learning rate (overrides PAR_FILE setting)")
    p.add_argument("--num-global-cond", "-gc", type=int, metavar="INT",
                   help="Number of global conditioning categories")
    # new (not in the reference): reproducible initialisation / shuffling
    p.add_argument("--seed", type=int, default=None, help="Seed for weight init and file shuffling")
    # positional arguments
    p.add_argument("ckpt_path", type=str, metavar="CKPT_PATH_PFX",
                   help="E.g. /path/to/ckpt/pfx, a path and prefix combination for writing checkpoint files")
    p.add_argument("arch_file", type=str, metavar="ARCH_FILE", help="JSON file specifying architectural parameters")
    p.add_argument("par_file", type=str, metavar="PAR_FILE",
                   help="JSON file specifying training and other hyperparameters")
    p.add_argument("sam_file", type=str, metavar="SAMPLES_FILE",
                   help="File containing lines:\n<id1>\\t/path/to/sample1.wav.npy\\t/path/to/sample1.mel.npy\n")
    return p.parse_args(argv)


def main(argv=None):
    args = get_args(argv)
    with open(args.arch_file, "r") as fp:
        arch = json.load(fp)
    with open(args.par_file, "r") as fp:
        par = json.load(fp)

    # args consistency checks (reference train.py:81-88)
    if args.num_global_cond is None and "n_gc_category" not in arch and arch.get("n_gc_embed", 0) > 0:
        print("Error: must provide n_gc_category in ARCH_FILE, or --num-global-cond", file=stderr)
        raise SystemExit(1)
    if args.tf_eager and args.tf_debug:
        print("Error: --tf-debug and --tf-eager cannot both be set", file=stderr)
        raise SystemExit(1)

    from lb_wavenet_b200 import config, data, tmodel
    from lb_wavenet_b200.dist import DistContext

    par = config.normalize_par(par)
    # overrides (reference train.py:110-120)
    if args.batch_size is not None:
        par["batch_sz"] = args.batch_size
    if args.slice_size is not None:
        par["slice_sz"] = args.slice_size
    if args.l2_factor is not None:
        par["l2_factor"] = args.l2_factor
    if args.learning_rate is not None:
        par["learning_rate"] = args.learning_rate

    ctx = DistContext.from_env()
    seed = args.seed
    if ctx.world > 1 and seed is None:
        seed = 0  # replicas must agree on the initial weights and the file order

    arch_n = config.normalize_arch(arch, None)
    mel_hop_sz = config.mel_hop_sz(arch_n)  # train.py:129-130

    dset_ckpt = "{}.dset".format(args.ckpt_path)
    dset = data.MaskedSliceWav(None, args.sam_file, par["sample_rate"], par["slice_sz"], par["prefetch_sz"],
                               arch_n["n_lc_in"], mel_hop_sz, par["batch_sz"], par["n_keep_checkpoints"], dset_ckpt,
                               args.resume_step or 0, dist=ctx, random_seed=seed,
                               wav_input_type=arch_n.get("wav_input_type", "mu_law_quant"))
    dset.init_sample_catalog()

    if args.num_global_cond is not None:  # train.py:140-146
        if args.num_global_cond < dset.get_max_id():
            print("Error: --num-global-cond must be >= {}, the highest ID in the dataset.".format(
                dset.get_max_id()), file=stderr)
            raise SystemExit(1)
        arch_n = config.normalize_arch(arch, args.num_global_cond, warn=False)
    elif arch_n["n_gc_embed"] > 0 and arch_n["n_gc_category"] < dset.get_max_id():
        print("Error: n_gc_category must be >= {}, the highest ID in the dataset.".format(dset.get_max_id()),
              file=stderr)
        raise SystemExit(1)

    net_ckpt = "{}.net".format(args.ckpt_path)
    net = tmodel.WaveNetTrain(
        **arch_n,
        batch_sz=par["batch_sz"],
        l2_factor=par["l2_factor"],
        add_summary=par["add_summary"],
        n_keep_checkpoints=par["n_keep_checkpoints"],
        ckpt_path=net_ckpt,
        resume_step=args.resume_step or 0,
        n_valid_total=par["n_valid_total"],
        sess=None,
        print_interval=args.progress_interval,
        dist=ctx, init_seed=seed)

    # the dataset depends on the net's receptive field (train.py:165-168)
    dset.set_receptive_field_size(net.get_recep_field_sz())
    dset.build()
    dset.init_vars()

    optimizer = tmodel.AdamOptimizer(learning_rate=par["learning_rate"])
    file_read_count, *data_ops = dset.get_op()
    grads_and_vars_op, loss_op = net.build(*data_ops)
    print("Built graph.", file=stderr)
    apply_grads_op = optimizer.apply_gradients(grads_and_vars_op)
    net.init_vars()

    if args.resume_step:
        net.restore()
        dset.restore()
        print("Restored net and dset from checkpoint", file=stderr)

    print("Starting training...", file=stderr)
    step = args.resume_step or 1
    loss = None
    while step < args.max_steps:
        if step == 5 and args.timeline_file is not None:
            _traced_step(net, apply_grads_op, loss_op, args.timeline_file)
        else:
            _, loss = net.run([apply_grads_op, loss_op])
        if step % args.save_interval == 0 and step != args.resume_step:
            net_save_path = net.save(step, write=ctx.rank == 0)  # every rank: SAVE rows are gathered
            dset_save_path = dset.save(step, net.file_read_count) if ctx.rank == 0 else None
            if ctx.rank == 0:
                print("Saved checkpoints to {} and {}".format(net_save_path, dset_save_path), file=stderr)
        step += 1
    dset._shutdown()
    return loss


def _traced_step(net, apply_grads_op, loss_op, path):
    """One normal step with per-category CUDA-event timing, dumped as Chrome-trace JSON."""
    import ctypes as C
    from lb_wavenet_b200 import _lib
    lib = _lib.load()
    lib.wn_prof_enable(1)
    net.run([apply_grads_op, loss_op])
    ms = (C.c_double * 16)()
    n = (C.c_int64 * 16)()
    lib.wn_prof_collect(ms, n)
    lib.wn_prof_enable(0)
    names = ["prep_embed_save", "layer_fwd", "post_fwd_loss", "post_bwd", "layer_bwd_gate", "layer_bwd_data",
             "wgrad", "pre_gc_bwd", "adam", "gen"]
    t, ev = 0.0, []
    for k, name in enumerate(names):
        if n[k]:
            ev.append(dict(name=name, ph="X", ts=t * 1e3, dur=ms[k] * 1e3, pid=0, tid=0, args=dict(launches=int(n[k]))))
            t += ms[k]
    with open(path, "w") as f:
        json.dump(dict(traceEvents=ev, displayTimeUnit="ms"), f)


if __name__ == "__main__":
    main()
