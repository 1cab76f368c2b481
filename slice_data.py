#!/usr/bin/env python
"""Offline slicer: same command line as the reference's slice_data.py (slice_data.py:47-66).

Cuts [beg, beg+sz) (both rounded DOWN to a multiple of the mel hop, slice_data.py:23-24) out of every
.wav.npy / .mel.npy pair of a catalog, skipping files that are too short (slice_data.py:8-11), and
writes a new catalog.  Plain host code: nothing here is on the GPU hot path.
"""
import argparse
import os
from sys import stderr

import numpy as np


def hop_floor(beg, sz, hop):
    return beg - beg % hop, sz - sz % hop


def read_catalog(path):
    rows = []
    with open(path) as fh:
        for line in fh:
            if line.strip():
                vid, wav_path, mel_path = line.rstrip("\n").split("\t")
                rows.append((int(vid), wav_path, mel_path))
    return rows


def sliced_name(src, sub_dir, out_dir):
    return "{}/{}/{}".format(out_dir, sub_dir, os.path.basename(src).replace(".npy", ".slice.npy"))


def slice_pair(wav_in, mel_in, wav_out, mel_out, hop, beg, sz):
    wav = np.load(wav_in)
    if len(wav) < beg + sz:
        print("Skipping {} of length {}".format(wav_in, len(wav)), file=stderr)
        return False
    mel = np.load(mel_in)
    assert len(mel) * hop == len(wav), "{}: mel frames * hop != wav length".format(wav_in)
    np.save(wav_out, wav[beg:beg + sz])
    np.save(mel_out, mel[beg // hop: beg // hop + sz // hop])
    return True


def get_args(argv=None):
    p = argparse.ArgumentParser(description="Slice Data")
    p.add_argument("--hop-size", "-hs", type=int, default=256, metavar="INT", help="Hop size of the Mel files")
    p.add_argument("--start-pos", "-sp", type=int, default=1024, metavar="INT", help="Start position for the slice")
    p.add_argument("--slice-size", "-ss", type=int, default=20480, metavar="INT", help="Size of the slice")
    p.add_argument("in_rdb_file", metavar="RDB_FILE", type=str,
                   help="File containing lines: <id>\\t/path/to/sample.wav.npy\\t/path/to/sample.mel.npy")
    p.add_argument("out_dir", metavar="OUT_DIR", type=str, help="Output directory for writing sliced files")
    p.add_argument("out_rdb_file", metavar="OUT_RDB_FILE", type=str, help="Name of the output rdb file")
    return p.parse_args(argv)


def main(argv=None):
    args = get_args(argv)
    beg, sz = hop_floor(args.start_pos, args.slice_size, args.hop_size)
    for d in (args.out_dir, args.out_dir + "/audio", args.out_dir + "/mel"):
        os.makedirs(d, exist_ok=True)
    with open(args.out_rdb_file, "w") as out:
        for vid, wav_path, mel_path in read_catalog(args.in_rdb_file):
            w_out = sliced_name(wav_path, "audio", args.out_dir)
            m_out = sliced_name(mel_path, "mel", args.out_dir)
            if slice_pair(wav_path, mel_path, w_out, m_out, args.hop_size, beg, sz):
                print("{}\t{}\t{}".format(vid, w_out, m_out), file=out)


if __name__ == "__main__":
    main()
