"""par/arch*.json and par/par*.json loading with the schema drift of the reference normalised.

The reference splats the arch dict into WaveNetTrain(**arch, ...) (train.py:152-153) and indexes
par[...] directly (train.py:133-136,154-160,178), but its shipped files drifted apart
(SURVEY.md section 8b): arch1/arch3 use ``n_post1`` and lack the LC / bias / input-type keys,
arch2/arch4 lack ``n_gc_category`` (given on the command line, train.py:81-84,140-146), arch4 has an
extra ``lc_hop_sz``, par3 says ``max_to_keep`` and lacks three keys.  All of them load here.
"""
from __future__ import annotations

import json
from sys import stderr
from typing import Optional

ARCH_DEFAULTS = dict(n_gc_embed=0, n_gc_category=None, n_lc_in=0, n_lc_out=0, lc_upsample=[],
                     use_bias=True, wav_input_type="mu_law_quant")
ARCH_REQUIRED = ("n_blocks", "n_block_layers", "n_quant", "n_res", "n_dil", "n_skip", "n_post")
PAR_DEFAULTS = dict(sample_rate=16000, prefetch_sz=2, add_summary=False, n_keep_checkpoints=10,
                    n_valid_total=1, l2_factor=0.0, learning_rate=1e-3)
PAR_REQUIRED = ("batch_sz", "slice_sz")


class ConfigError(ValueError):
    pass


def normalize_arch(arch: dict, num_global_cond: Optional[int] = None, warn: bool = True) -> dict:
    """Return the key set WaveNetTrain.__init__ consumes (reference tmodel.py:8-24)."""
    a = dict(arch)
    if "n_post" not in a and "n_post1" in a:  # generate.py:61-71 spelling
        a["n_post"] = a.pop("n_post1")
    a.pop("n_post1", None)
    a.pop("lc_hop_sz", None)  # par/arch4.json:12, consumed by nothing
    for k, v in ARCH_DEFAULTS.items():
        if k not in a:
            if k == "use_bias" and warn:
                print("Warning: arch file has no 'use_bias'; defaulting to true", file=stderr)
            a[k] = v
    if num_global_cond is not None:  # train.py:140-146
        a["n_gc_category"] = num_global_cond
    if a["n_gc_category"] is None:
        if a["n_gc_embed"] > 0:
            raise ConfigError("must provide n_gc_category in ARCH_FILE, or --num-global-cond")  # train.py:81-84
        a["n_gc_category"] = 0
    missing = [k for k in ARCH_REQUIRED if k not in a]
    if missing:
        raise ConfigError("arch file lacks keys: %s" % ", ".join(missing))
    if a["wav_input_type"] not in ("mu_law_quant", "raw"):
        raise ConfigError("wav_input_type must be 'mu_law_quant' or 'raw' (tmodel.py:59-62)")
    return a


def normalize_par(par: dict) -> dict:
    p = dict(par)
    if "n_keep_checkpoints" not in p and "max_to_keep" in p:  # par/par3.json
        p["n_keep_checkpoints"] = p.pop("max_to_keep")
    p.pop("max_to_keep", None)
    for k, v in PAR_DEFAULTS.items():
        p.setdefault(k, v)
    missing = [k for k in PAR_REQUIRED if k not in p]
    if missing:
        raise ConfigError("par file lacks keys: %s" % ", ".join(missing))
    return p


def load_arch(path: str, num_global_cond: Optional[int] = None) -> dict:
    with open(path, "r") as fp:
        return normalize_arch(json.load(fp), num_global_cond)


def load_par(path: str) -> dict:
    with open(path, "r") as fp:
        return normalize_par(json.load(fp))


def mel_hop_sz(arch: dict) -> int:
    """reference train.py:129-130: product of lc_upsample (1 when there is no LC stack)."""
    hop = 1
    for s in arch.get("lc_upsample", []) or []:
        hop *= int(s)
    return hop


def _pad_res_dil(n: int) -> int:
    """The tcgen05 layer kernels take n_res = n_dil = 32 (fused per-layer kernels) or multiples of 64 (GEMM chain)."""
    return 32 if n <= 32 else 64 if n <= 64 else (n + 63) // 64 * 64


def _pad_skip_post(n: int) -> int:
    return max(64, (n + 63) // 64 * 64)


def engine_arch(arch: dict, pad: bool = True) -> dict:
    """The keys the C ABI's wn_arch carries.  Channel counts the kernels do not tile (the reference's par/arch2.json has
    n_res = 3, n_dil = 4, n_skip = 8, n_post = 6) are PADDED here: the device arena holds zero-extended tensors, the
    host mirror (tmodel / imodel) reads and writes the logical slices, so checkpoints keep the reference's shapes.  Zero
    padding is closed under the whole training step: a padded input channel is 0, a padded output channel gets
    tanh(0) * sigmoid(0) = 0 / relu(0) = 0, every gradient into a padded row or column is a product with one of those
    zeros, and Adam / L2 leave an exactly-zero weight with an exactly-zero gradient at zero."""
    out = dict(n_blocks=arch["n_blocks"], n_block_layers=arch["n_block_layers"], n_quant=arch["n_quant"],
               n_res=arch["n_res"], n_dil=arch["n_dil"], n_skip=arch["n_skip"], n_post=arch["n_post"],
               n_gc_embed=arch["n_gc_embed"], n_gc_category=arch["n_gc_category"],
               use_bias=1 if arch["use_bias"] else 0)
    if arch.get("n_lc_out", 0):  # local conditioning (tmodel.py:15-17): mel channels, conditioning channels, strides
        out.update(n_lc_in=int(arch["n_lc_in"]), n_lc_out=int(arch["n_lc_out"]),
                   lc_upsample=[int(x) for x in arch["lc_upsample"]])
    if pad:
        out["n_res"], out["n_dil"] = _pad_res_dil(out["n_res"]), _pad_res_dil(out["n_dil"])
        if out["n_res"] != out["n_dil"] and max(out["n_res"], out["n_dil"]) <= 64:
            out["n_res"] = out["n_dil"] = max(out["n_res"], out["n_dil"])   # the fused kernels want n_res == n_dil
        out["n_skip"], out["n_post"] = _pad_skip_post(out["n_skip"]), _pad_skip_post(out["n_post"])
    return out
