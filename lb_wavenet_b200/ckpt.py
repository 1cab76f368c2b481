"""Checkpoint base class with the reference's file naming and key layout.

Mirrors reference ckpt.py:13-81: a dict of saveable objects keyed by serial name, saved to
``<ckpt_path>-<step>.{index,meta,data-00000-of-00001}`` keeping the newest
``n_keep_checkpoints`` (tf.train.Saver max_to_keep, ckpt.py:41-42), restored from
``<ckpt_path>-<resume_step>`` after checking that the three files are readable (ckpt.py:70-76).

Container: TensorFlow's own tensor-bundle format, written and read without TensorFlow by ``tfbundle.py``
(``.index`` = SSTable of BundleHeaderProto / BundleEntryProto records, ``.data-00000-of-00001`` = raw little-endian
tensor bytes in key order), so checkpoints interchange with reference-trained ones; ``.meta`` (a MetaGraphDef in the
reference, of which only the presence is ever checked, ckpt.py:70-76) is a small stub.  Checkpoints written by
earlier builds of this repository (JSON ``.index``) are still readable.
"""
from __future__ import annotations

import json
import os
import re
import zlib
from sys import stderr
from typing import Callable, Dict, List, Optional

import numpy as np

SUFFIXES = ("index", "meta", "data-00000-of-00001")


def _expand_ckpt(ckpt: str) -> List[str]:
    """reference ckpt.py:8-11"""
    return ["{}.{}".format(ckpt, s) for s in SUFFIXES]


class Variable:
    """A saveable object: name, dtype/shape, and accessors into wherever the value lives
    (a view of the device arena, a host scalar, ...)."""

    def __init__(self, name: str, shape, dtype, get: Callable[[], np.ndarray],
                 set: Callable[[np.ndarray], None], trainable: bool = True, optional: bool = False):
        self.name, self.shape, self.dtype = name, tuple(shape), np.dtype(dtype)
        self._get, self._set, self.trainable = get, set, trainable
        # optional: an extra key beyond the reference's layout (optimiser slots, exact loader state).  Written always,
        # and simply skipped on restore when a checkpoint (e.g. one trained by the reference) does not have it.
        self.optional = optional

    def numpy(self) -> np.ndarray:
        # (np.ascontiguousarray would turn a 0-d scalar into shape (1,))
        return np.array(np.asarray(self._get()), dtype=self.dtype, order="C").reshape(self.shape)

    def assign(self, value) -> None:
        v = np.asarray(value)
        if tuple(v.shape) != self.shape:
            raise ValueError("shape mismatch restoring {}: checkpoint {} vs variable {}".format(
                self.name, tuple(v.shape), self.shape))
        self._set(v.astype(self.dtype))


class Checkpoint(object):
    def __init__(self, ckpt_path, n_keep_checkpoints, resume_step, sess=None):
        self.n_keep_checkpoints = n_keep_checkpoints
        self.ckpt_path = ckpt_path
        self.resume_step = resume_step
        self.sess = sess  # accepted for signature compatibility, unused
        self.saveable_objects: Dict[str, Variable] = {}
        self.initializable_ops: list = []
        self.initialized = False

    # reference ckpt.py:26-31
    def add_saveable_objects(self, objs: Dict[str, Variable]):
        self.saveable_objects.update(objs)
        self.initialized = True

    def add_initializable_ops(self, ops):
        self.initializable_ops += list(ops)

    def init_vars(self):
        """reference ckpt.py:44-51: run the initialisers (callables here)."""
        for op in self.initializable_ops:
            op()

    def _all_saved_steps(self) -> List[int]:
        d = os.path.dirname(self.ckpt_path) or "."
        base = os.path.basename(self.ckpt_path)
        pat = re.compile(re.escape(base) + r"-(\d+)\.index$")
        steps = []
        if os.path.isdir(d):
            for fn in os.listdir(d):
                m = pat.match(fn)
                if m:
                    steps.append(int(m.group(1)))
        return sorted(steps)

    def save(self, step: int, write: bool = True) -> str:
        """reference ckpt.py:54-62 -> '<ckpt_path>-<step>'.  Data-parallel ranks all call save() (reading a
        sharded variable is a collective gather); only the rank with write=True touches the disk."""
        if not self.initialized:
            raise ValueError("add_saveable_objects has not been called")
        path_pfx = "{}-{}".format(self.ckpt_path, int(step))
        values = {key: self.saveable_objects[key].numpy() for key in sorted(self.saveable_objects)}
        if not write:
            return path_pfx
        d = os.path.dirname(path_pfx)
        if d:
            os.makedirs(d, exist_ok=True)
        from . import tfbundle
        tfbundle.write_bundle(path_pfx, values, suffix=".tmp")
        with open(path_pfx + ".meta.tmp", "wb") as f:  # presence only (ckpt.py:70-76); not a real MetaGraphDef
            f.write(b"\x0a\x1c\x0a\x1alb-wavenet-b200 (no graph)")
        for s in SUFFIXES:
            os.replace(path_pfx + "." + s + ".tmp", path_pfx + "." + s)
        # TF 'checkpoint' state file + max_to_keep pruning (ckpt.py:41-42)
        steps = self._all_saved_steps()
        keep = self.n_keep_checkpoints if self.n_keep_checkpoints and self.n_keep_checkpoints > 0 else len(steps)
        for old in steps[:-keep] if len(steps) > keep else []:
            for fn in _expand_ckpt("{}-{}".format(self.ckpt_path, old)):
                try:
                    os.remove(fn)
                except OSError:
                    pass
        steps = steps[-keep:]
        with open(os.path.join(d or ".", "checkpoint"), "w") as f:
            base = os.path.basename(self.ckpt_path)
            f.write('model_checkpoint_path: "{}-{}"\n'.format(base, int(step)))
            for s in steps:
                f.write('all_model_checkpoint_paths: "{}-{}"\n'.format(base, s))
        return path_pfx

    def restore(self, ckpt_file: Optional[str] = None):
        """reference ckpt.py:65-81"""
        from os import access, R_OK
        if ckpt_file is None:
            ckpt_file = "{}-{}".format(self.ckpt_path, self.resume_step)
        print("Restoring from {}".format(ckpt_file))
        for fn in _expand_ckpt(ckpt_file):
            if not access(fn, R_OK):
                print("Couldn't find checkpoint file {}".format(fn), file=stderr)
                raise SystemExit(1)
        tensors = read_checkpoint(ckpt_file)
        missing = [k for k, v in self.saveable_objects.items() if k not in tensors and not v.optional]
        if missing:
            raise KeyError("checkpoint {} lacks keys: {}".format(ckpt_file, ", ".join(missing[:8])))
        self.restored_optional = []
        for k, var in self.saveable_objects.items():
            if k in tensors:
                var.assign(tensors[k])
                if var.optional:
                    self.restored_optional.append(k)


def read_checkpoint(ckpt_file: str) -> Dict[str, np.ndarray]:
    """All tensors of a checkpoint prefix, keyed by serial name (TF tensor bundle, or this repository's first JSON
    container)."""
    from . import tfbundle
    if tfbundle.is_table_file(ckpt_file + ".index"):
        return tfbundle.read_bundle(ckpt_file)
    with open(ckpt_file + ".index", "r") as f:
        idx = json.load(f)
    out = {}
    with open(ckpt_file + ".data-00000-of-00001", "rb") as f:
        blob = f.read()
    for k, e in idx["tensors"].items():
        raw = blob[e["offset"]:e["offset"] + e["size"]]
        if (zlib.crc32(raw) & 0xFFFFFFFF) != e["crc32"]:
            raise IOError("checkpoint {}: crc mismatch for {}".format(ckpt_file, k))
        out[k] = np.frombuffer(raw, dtype=np.dtype(e["dtype"]).newbyteorder("<")).reshape(e["shape"]).copy()
    return out
