"""Shared helpers for the parity tests (oracle side)."""
import numpy as np
import torch

from oracle import wavenet_oracle as O

TINY = dict(n_blocks=2, n_block_layers=4, n_quant=256, n_res=32, n_dil=32, n_skip=64, n_post=64,
            n_gc_embed=0, n_gc_category=0, use_bias=1)
TINY_GC = dict(TINY, n_gc_embed=17, n_gc_category=11)
TINY_ASYM = dict(n_blocks=1, n_block_layers=5, n_quant=256, n_res=16, n_dil=48, n_skip=80, n_post=32,
                 n_gc_embed=0, n_gc_category=0, use_bias=1)
TINY_NOBIAS = dict(TINY, use_bias=0)
WIDE = dict(n_blocks=1, n_block_layers=3, n_quant=256, n_res=128, n_dil=128, n_skip=512, n_post=512,
            n_gc_embed=0, n_gc_category=0, use_bias=1)
WIDE_DEEP = dict(WIDE, n_block_layers=10)  # dilations 1..512 through the wide-layer GEMM kernels
CLASSIC = dict(n_blocks=3, n_block_layers=10, n_quant=256, n_res=32, n_dil=32, n_skip=256, n_post=256,
               n_gc_embed=0, n_gc_category=0, use_bias=1)
CLASSIC_SHALLOW = dict(CLASSIC, n_blocks=1, n_block_layers=4)
C1 = dict(n_blocks=5, n_block_layers=10, n_quant=256, n_res=32, n_dil=32, n_skip=512, n_post=512,
          n_gc_embed=17, n_gc_category=377, use_bias=1)


# local conditioning (reference tmodel.py:68-83,156-160): small strides / channel counts, and the reference's own
# par/arch5.json -- the one shipped architecture that satisfies train.py as written
TINY_LC = dict(TINY, n_lc_in=20, n_lc_out=24, lc_upsample=[2, 4])
TINY_GC_LC = dict(TINY_GC, n_lc_in=12, n_lc_out=16, lc_upsample=[4])
ARCH5 = dict(n_blocks=5, n_block_layers=10, n_quant=256, n_res=32, n_dil=32, n_skip=512, n_post=512,
             n_gc_embed=16, n_gc_category=376, use_bias=1, n_lc_in=80, n_lc_out=80, lc_upsample=[4, 4, 4, 4])


def oracle_arch(d):
    return O.Arch(d["n_blocks"], d["n_block_layers"], d["n_quant"], d["n_res"], d["n_dil"], d["n_skip"],
                  d["n_post"], d["n_gc_embed"], d["n_gc_category"], bool(d["use_bias"]),
                  n_lc_in=d.get("n_lc_in", 0), n_lc_out=d.get("n_lc_out", 0),
                  lc_upsample=tuple(d.get("lc_upsample", ())))


def synth_mel(B, T, a, seed, wav=None):
    """mel frames [B, T / hop, n_lc_in] for an architecture with local conditioning (None otherwise).  With `wav`
    (mu-law codes [B, T]) the frames are a cosine-transform spectrum of the very samples they condition, plus a little
    noise -- like the real data, where mel.npy is computed from wav.npy.  That matters for gradient tests: with mel
    independent of the audio the LC gradients have no coherent part, they are a random sum whose relative rounding error
    equals the per-element error of dv and does not fall with the number of positions (measured with the CPU oracle,
    8 192 positions: 6.7 % with random mel, 2.7 % with correlated mel, other tensors 2.1-2.4 % either way)."""
    if not a.has_lc():
        return None
    rng = np.random.default_rng(seed)
    hop, n = a.lc_hop(), a.n_lc_in
    if wav is None:
        return rng.normal(0, 1.0, (B, T // hop, n)).astype(np.float32)
    x = O.mu_decode_np(np.asarray(wav, np.int64).clip(0, 255)).reshape(B, T // hop, hop).astype(np.float64)
    t = np.arange(hop) + 0.5
    basis = np.cos(np.pi * (np.arange(n)[:, None] + 0.5) * t[None, :] / hop)
    return (4.0 * x @ basis.T / np.sqrt(hop) + 0.1 * rng.normal(size=(B, T // hop, n))).astype(np.float32)


def synth_batch(B, T, n_cat, seed, invalid_frac=0.3):
    """Synthetic mu-law codes + id masks with file junctions at random places."""
    rng = np.random.default_rng(seed)
    t = np.arange(T)[None, :] + rng.integers(0, 1000, (B, 1))
    sig = 0.5 * np.sin(t * 0.05) + 0.3 * np.sin(t * 0.31 + 1.0) + 0.05 * rng.normal(size=(B, T))
    wav = O.mu_encode_np(np.clip(sig, -1, 1).astype(np.float32)).astype(np.int32)
    ids = np.empty((B, T), np.int32)
    for b in range(B):
        pos = 0
        while pos < T:
            seg = int(rng.integers(max(2, T // 6), max(3, T // 2)))
            vid = int(rng.integers(1, max(2, n_cat + 1)))
            inval = int(seg * invalid_frac)
            ids[b, pos:pos + inval] = 0
            ids[b, pos + inval:pos + seg] = vid
            pos += seg
    return wav, ids


def scaled_params(a, B, seed, scale=1.0, bias_scale=0.2):
    """Xavier init (reference arch.py:63) with non-zero biases; SAVE = bf16-representable noise."""
    p = O.init_params(a, B, seed=seed, bias_scale=bias_scale)
    for k in p:
        if k.startswith("SAVE"):
            p[k] = torch.tensor(p[k] * 20).to(torch.bfloat16).float().numpy()
        elif p[k].dtype.kind == "f":
            p[k] = (p[k] * scale).astype(np.float32)
    return p


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def record(key, value):
    """Append a measured parity figure to gpurun_out/parity_measured.jsonl (best effort)."""
    import json
    import os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_measured.jsonl"), "a") as f:
            f.write(json.dumps({key: value}) + "\n")
    except OSError:
        pass
