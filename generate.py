#!/usr/bin/env python
"""Generation entry point: same command line as the reference's generate.py (generate.py:5-29).

    generate.py [-w WAV] [-ts F] [-td F] [-g F=5] [-s INT=16000] [-c INT=1000] [-b INT=10]
                ARCH_FILE CHECKPOINT_PREFIX OUTPUT_WAV_DIR

Writes OUTPUT_WAV_DIR/gen.i{n}.wav for every stream (generate.py:112-116).  New options:
--seed (sampling seed; the reference's tf.multinomial is unseeded) and --gc-ids (voice id per
stream; the reference hard-codes [5, 6], generate.py:94).
"""
import argparse
import json
import os
from sys import stderr


def get_args(argv=None):
    p = argparse.ArgumentParser(description="WaveNet")
    p.add_argument("--teacher-wav", "-w", type=str,
                   help="Provide a preliminary teacher-forcing vector to prime the generation")
    p.add_argument("--teacher-start", "-ts", type=float, help="Number of seconds to skip in <teacher_wav>")
    p.add_argument("--teacher-duration", "-td", type=float, help="Number of seconds to parse from <teacher_wav>")
    p.add_argument("--gen-seconds", "-g", type=float, default=5, help="Number of additional seconds to generate")
    p.add_argument("--sample-rate", "-s", type=int, default=16000,
                   help="Number of samples per second for parsed .wav files")
    p.add_argument("--chunk-size", "-c", type=int, default=1000,
                   help="Number of timesteps generated per kernel launch (the reference's buffer-shift interval)")
    p.add_argument("--batch-size", "-b", type=int, default=10, help="Number of .wav files to generate simultaneously")
    p.add_argument("--seed", type=int, default=None, help="Sampling seed (default: fresh entropy)")
    p.add_argument("--gc-ids", type=str, default=None, help="Comma separated voice ids, one per stream")
    p.add_argument("arch_file", type=str, metavar="ARCH_FILE", help="JSON file specifying architectural parameters")
    p.add_argument("ckpt", metavar="CHECKPOINT_PREFIX", type=str,
                   help="Provide <ckpt> for <ckpt>.{meta,index,data-..} files")
    p.add_argument("wav_dir", metavar="OUTPUT_WAV_DIR", type=str, help="Output directory for generated .wav files")
    return p.parse_args(argv)


def main(argv=None):
    args = get_args(argv)
    from lb_wavenet_b200 import ckpt, config, imodel, wavio

    with open(args.arch_file, "r") as fp:
        arch = config.normalize_arch(json.load(fp), None, warn=False)
    for fn in ckpt._expand_ckpt(args.ckpt):  # generate.py:47-50
        if not os.access(fn, os.R_OK):
            print("Couldn't find checkpoint file {}".format(fn), file=stderr)
            raise SystemExit(1)

    teacher_vec, teacher_seconds = None, 0
    if args.teacher_wav is not None:  # generate.py:52-59
        teacher_vec = wavio.read_wav(args.teacher_wav, args.sample_rate, args.teacher_start, args.teacher_duration)
        teacher_seconds = teacher_vec.shape[0] / args.sample_rate

    net = imodel.WaveNetGen(arch["n_blocks"], arch["n_block_layers"], arch["n_quant"], arch["n_res"], arch["n_dil"],
                            arch["n_skip"], arch["n_post"], arch["n_gc_embed"], arch["n_gc_category"],
                            arch["use_bias"], args.batch_size, args.chunk_size, teacher_vec, seed=args.seed)
    print("Building graph.")
    wave_ops = net.build_graph()
    print("Restoring from {}".format(args.ckpt))
    net.restore(None, args.ckpt)
    print("Initializing buffers.")
    net.init_buffers(None)

    gen_sz = int((args.gen_seconds + teacher_seconds) * args.sample_rate)
    feed_dict = {net.gen_sz: gen_sz}
    if net.use_gc:
        if args.gc_ids is not None:
            ids = [int(v) for v in args.gc_ids.split(",")]
        else:
            ids = [1 + (i % max(1, arch["n_gc_category"])) for i in range(args.batch_size)]
        feed_dict[net.gc_ids] = ids

    print("Starting inference...")
    n, wav_streams, wpos = net.run(feed_dict)
    wav_streams = wav_streams.cpu().numpy()

    print("Writing wav files.")
    os.makedirs(args.wav_dir, exist_ok=True)
    for i in range(wav_streams.shape[0]):
        path = os.path.join(args.wav_dir, "gen.i{}.wav".format(i))
        wavio.write_wav(path, wav_streams[i], args.sample_rate)
        print("Wrote {}".format(path))
    print("Finished.")


if __name__ == "__main__":
    main()
