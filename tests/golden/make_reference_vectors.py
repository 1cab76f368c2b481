#!/usr/bin/env python
"""Golden vectors produced by the REFERENCE'S OWN PYTHON, run in this container.

    python tests/golden/make_reference_vectors.py        (needs /root/reference; writes tests/golden/reference_*.npz)

hrbigelow/lb-wavenet is TensorFlow-1.x graph code and TensorFlow cannot be installed here, so the reference's modules
(tmodel.py, arch.py, ops.py, ckpt.py, data.py) are imported UNMODIFIED from /root/reference with oracle/tf1_shim on
sys.path: a ~400-line eager stand-in that carries out every `tf.*` call the reference makes on torch CPU tensors
(float64; see its docstring for what that does and does not pin).  The vectors written here are what
`WaveNetTrain.build()` (tmodel.py:284-339), `ops.mu_encode / mu_decode` (ops.py:4-39) and
`MaskedSliceWav._gen_slice_batch` (data.py:110-227) return; tests/test_reference_vectors.py holds the oracle -- and, on
the GPU, the CUDA path -- to them.  Inputs and parameters are regenerated from seeds by `cases()` below, which the
tests import, so only outputs are stored.
"""
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

# ---- the cases: everything the tests need to rebuild the inputs ---------------------------------------------------
TRAIN_CASES = {
    # name: (arch dict in the reference's arch.json vocabulary, batch, slice, l2_factor, seed)
    "r32": (dict(n_blocks=2, n_block_layers=4, n_quant=256, n_res=32, n_dil=32, n_skip=256, n_post=256, n_gc_embed=0,
                 n_gc_category=0, n_lc_in=0, n_lc_out=0, lc_upsample=[], use_bias=True), 2, 48, 1e-3, 101),
    "gc": (dict(n_blocks=2, n_block_layers=3, n_quant=256, n_res=32, n_dil=32, n_skip=64, n_post=64, n_gc_embed=8,
                n_gc_category=3, n_lc_in=0, n_lc_out=0, lc_upsample=[], use_bias=True), 3, 40, 1e-3, 102),
    "lc": (dict(n_blocks=1, n_block_layers=4, n_quant=256, n_res=32, n_dil=32, n_skip=64, n_post=64, n_gc_embed=0,
                n_gc_category=0, n_lc_in=5, n_lc_out=6, lc_upsample=[2, 3], use_bias=True), 2, 48, 0.0, 103),
    # the reference's par/arch2.json channel counts, no biases
    "odd": (dict(n_blocks=2, n_block_layers=3, n_quant=256, n_res=3, n_dil=4, n_skip=8, n_post=6, n_gc_embed=0,
                 n_gc_category=0, n_lc_in=0, n_lc_out=0, lc_upsample=[], use_bias=False), 2, 32, 1e-2, 104),
    # the reference's par/arch5.json as shipped -- the one architecture file train.py accepts as written: 5 x 10 layers,
    # S = P = 512, 16-wide voice embedding over 376 categories, 80 -> 80 local conditioning upsampled 4 x 4 x 4 x 4 = 256
    "arch5": (dict(n_blocks=5, n_block_layers=10, n_quant=256, n_res=32, n_dil=32, n_skip=512, n_post=512, n_gc_embed=16,
                   n_gc_category=376, n_lc_in=80, n_lc_out=80, lc_upsample=[4, 4, 4, 4], use_bias=True), 2, 256, 1e-4, 105),
}
RAW_CASES = {"raw"}  # wav_input_type 'raw': the graph receives float audio and encodes it itself (tmodel.py:59-62, ops.py:4-9)
TRAIN_CASES["raw"] = (dict(TRAIN_CASES["odd"][0], use_bias=True), 2, 32, 1e-2, 106)
N_STAGES = 2
SAMPLE_LIMIT = {"arch5": 48}  # elements stored per gradient tensor (default 256): arch5 has 760 variables
LOGIT_STRIDE = {"arch5": 16}  # every k-th timestep of the logits is stored (default 4)


def save_sample(x):
    """what is stored of a SAVE variable: all of it up to 4096 elements, else a strided sample of 512"""
    x = np.asarray(x)
    return x if x.size <= 4096 else sample_of(x, 512)


def sample_limit(name):
    return SAMPLE_LIMIT.get(name, 256)


def logit_stride(name):
    return LOGIT_STRIDE.get(name, 4)


def train_inputs(name, stage):
    """(wav int32 [B, T], ids int32 [B, T], mel float32 [B, T / hop, n_lc_in] or None) of one stage"""
    arch, B, T, _, seed = TRAIN_CASES[name]
    rng = np.random.default_rng(1000 * seed + stage)
    wav = rng.integers(0, arch["n_quant"], (B, T)).astype(np.int32)
    n_cat = max(1, arch["n_gc_category"])
    ids = np.zeros((B, T), np.int32)
    for b in range(B):  # a junction per slot: a run of invalid positions (id 0) followed by another voice
        j = int(rng.integers(T // 4, T // 2))
        ids[b, :j] = rng.integers(1, n_cat + 1)
        ids[b, j + 5:] = rng.integers(1, n_cat + 1)
    if stage == 0:
        ids[0, :7] = 0  # the start of a file: receptive-field positions are invalid
    wav[0, 9] = -1  # an out-of-range code at a valid position: tf.one_hot gives an all-zero row (input AND label)
    mel = None
    if arch["n_lc_out"] > 0:
        hop = int(np.prod(arch["lc_upsample"]))
        mel = rng.standard_normal((B, T // hop, arch["n_lc_in"])).astype(np.float32)
    if name in RAW_CASES:
        wav[0, 9] = 7  # (float audio has no out-of-range code)
    return wav, ids, mel


_RAW_TABLE = None


def raw_audio(codes):
    """float32 audio whose mu-law code is `codes`: the middle of the code's interval of the encoder (between two
    thresholds of ops.py:4-9), so that encoding it in float32 (TensorFlow) or float64 (the shim's default) gives the same code"""
    global _RAW_TABLE
    if _RAW_TABLE is None:
        from oracle import wavenet_oracle as O
        thr = O.mu_encode_thresholds(256).astype(np.float64)
        edges = np.concatenate([[-1.0], thr, [1.0]])
        _RAW_TABLE = (0.5 * (edges[:-1] + edges[1:])).astype(np.float32)
        assert np.array_equal(O.mu_encode_np(_RAW_TABLE), np.arange(256))
    return _RAW_TABLE[np.asarray(codes)]


def train_params(name):
    """name -> float32 array, keyed by the reference's serial names (arch.py:142)"""
    from oracle import wavenet_oracle as O
    arch, B, _, _, seed = TRAIN_CASES[name]
    a = oracle_arch(arch)
    return O.init_params(a, B, seed=seed, bias_scale=0.2 if arch["use_bias"] else 0.0)


def oracle_arch(arch):
    from oracle import wavenet_oracle as O
    return O.Arch(arch["n_blocks"], arch["n_block_layers"], arch["n_quant"], arch["n_res"], arch["n_dil"],
                  arch["n_skip"], arch["n_post"], arch["n_gc_embed"], arch["n_gc_category"], bool(arch["use_bias"]),
                  arch["n_lc_in"], arch["n_lc_out"], tuple(arch["lc_upsample"]))


def sample_of(x, limit=256):
    """what is stored of a large tensor: every k-th element of the flattened array"""
    f = np.asarray(x).reshape(-1)
    k = max(1, -(-f.size // limit))
    return f[::k]


DEAL_CASE = dict(batch_sz=3, slice_sz=50, mel_hop_sz=5, mel_spectrum_sz=4, recep_field_sz=12, n_batches=9,
                 file_lens=[133, 61, 9, 240, 77, 55, 102], vids=[3, 1, 2, 5, 4, 1, 2], n_epochs=3, seed=7)


def deal_files(tmpdir=None):
    """[(vid, wav int array, mel float array)] of the dealer case; written as .npy files under tmpdir when given"""
    c = DEAL_CASE
    rng = np.random.default_rng(c["seed"])
    out = []
    for i, (n, vid) in enumerate(zip(c["file_lens"], c["vids"])):
        wav = rng.integers(0, 256, n).astype(np.int32)
        mel = rng.standard_normal((n // c["mel_hop_sz"], c["mel_spectrum_sz"])).astype(np.float32)
        out.append((vid, wav, mel))
        if tmpdir is not None:
            np.save(os.path.join(tmpdir, "w%d.npy" % i), wav)
            np.save(os.path.join(tmpdir, "m%d.npy" % i), mel)
    return out


def deal_order():
    """file order the dealer case is fed (TF's shuffle itself is not reproducible outside TF: data.py:246-250)"""
    c = DEAL_CASE
    rng = np.random.default_rng(c["seed"] + 1)
    return np.concatenate([rng.permutation(len(c["file_lens"])) for _ in range(c["n_epochs"])])


def mu_inputs():
    rng = np.random.default_rng(5)
    pcm = np.arange(-32768, 32768, dtype=np.int32).astype(np.float32) / np.float32(32768.0)
    return np.concatenate([pcm, rng.uniform(-1, 1, 20000).astype(np.float32), np.float32([-1.0, 1.0, 0.0, -0.0])])


# ---- running the reference --------------------------------------------------------------------------------------
def _import_reference():
    assert os.path.isdir(REF), "the reference tree is needed to (re)generate the vectors"
    sys.path.insert(0, os.path.join(ROOT, "oracle", "tf1_shim"))
    sys.path.insert(0, REF)
    sys.modules.setdefault("librosa", types.ModuleType("librosa"))  # imported by data.py, never called on this path
    if not hasattr(np, "float"):
        np.float = float  # data.py:127 was written against numpy < 1.24
    import tensorflow as tf
    assert "tf1_shim" in tf.__file__
    import tmodel  # noqa: F401  (the reference's own modules)
    import ops
    import data
    return tf, tmodel, ops, data


def run_train_case(tf, tmodel, name):
    import torch
    arch, B, T, l2, _ = TRAIN_CASES[name]
    tf._reset()
    tf._set_float(torch.float64)
    tf._set_eager(False)
    net = tmodel.WaveNetTrain(**arch, wav_input_type="raw" if name in RAW_CASES else "mu_law_quant", batch_sz=B, l2_factor=l2, add_summary=False,
                              n_keep_checkpoints=1, ckpt_path="/tmp/none", resume_step=0, n_valid_total=10 ** 6,
                              sess=None, print_interval=10 ** 6)
    captured = {}
    post = net._postprocess

    def post_spy(x):  # the logits never leave build(); this only records what the reference's own method returns
        out = post(x)
        captured["logits"] = out[0].detach().numpy().copy()
        return out
    net._postprocess = post_spy

    def call_build(stage):
        wav, ids, mel = train_inputs(name, stage)
        lc = None if mel is None else torch.as_tensor(mel, dtype=torch.float64)
        if name in RAW_CASES:
            return net.build(torch.as_tensor(raw_audio(wav), dtype=torch.float64), lc, torch.as_tensor(ids))
        return net.build(torch.as_tensor(wav), lc, torch.as_tensor(ids))

    call_build(0)  # creates the variables (Xavier-initialised by the reference's own get_variable wrapper) ...
    p = train_params(name)
    assert set(net.vars.keys()) == set(p.keys()), sorted(set(net.vars.keys()) ^ set(p.keys()))
    for k, v in net.vars.items():  # ... which then receive the case's parameters under the reference's serial names
        assert tuple(v.shape) == tuple(np.shape(p[k])), (k, v.shape, np.shape(p[k]))
        v.load(p[k])
    out = {"var_order": np.array(list(net.vars.keys()))}
    for stage in range(N_STAGES):
        grads_vars, loss = call_build(stage)
        names = {id(v): k for k, v in net.vars.items()}
        out["s%d_loss" % stage] = np.float64(loss.detach().numpy())
        out["s%d_logits" % stage] = captured["logits"][:, ::logit_stride(name), :].astype(np.float32)
        out["s%d_global_step" % stage] = net.vars["GLOBAL_STEP"].numpy().copy()
        out["s%d_valid_samples" % stage] = net.vars["VALID_SAMPLES"].numpy().copy()
        for g, v in grads_vars:
            k = names[id(v)]
            g = np.zeros(v.shape) if g is None else g.detach().numpy()
            out["s%d_grad_%s" % (stage, k)] = sample_of(g, sample_limit(name)).astype(np.float64)
            out["s%d_gradnorm_%s" % (stage, k)] = np.float64(np.sqrt((g ** 2).sum()))
        for k, v in net.vars.items():
            if k.startswith("SAVE"):
                out["s%d_%s" % (stage, k)] = save_sample(v.numpy().astype(np.float64)).copy()
    return out


def run_mu(tf, ops):
    import torch
    tf._set_float(torch.float32)  # ops.py computes in float32 (tf.to_float)
    x = mu_inputs()
    codes = ops.mu_encode(torch.as_tensor(x), 256).numpy().astype(np.int32)
    dec = ops.mu_decode(torch.arange(256, dtype=torch.int32), 256).numpy().astype(np.float32)
    codes_np = ops.mu_encode_np(x, 256)
    dec_np = ops.mu_decode_np(np.arange(256, dtype=np.int32), 256).astype(np.float32)
    tf._set_float(torch.float64)
    return dict(codes_tf=codes, decoded_tf=dec, codes_np=codes_np.astype(np.int32), decoded_np=dec_np)


def run_dealer(tf, data):
    c = DEAL_CASE
    tf._reset()
    tf._set_eager(True)  # data.py:215: the eager file reader (the graph one needs a session)
    with tempfile.TemporaryDirectory() as tmp:
        deal_files(tmp)
        sam = os.path.join(tmp, "sam.txt")
        with open(sam, "w") as f:
            for i, vid in enumerate(c["vids"]):
                f.write("%d\t%s\t%s\n" % (vid, os.path.join(tmp, "w%d.npy" % i), os.path.join(tmp, "m%d.npy" % i)))
        ds = data.MaskedSliceWav(None, sam, 16000, c["slice_sz"], 1, c["mel_spectrum_sz"], c["mel_hop_sz"],
                                 c["batch_sz"], 1, "/tmp/none", 0)
        ds.init_sample_catalog()
        ds.set_receptive_field_size(c["recep_field_sz"])
        paths = list(ds._gen_path())  # the reference's own (vid, wav_path, mel_path) records

        class Item:  # what an eager tf.data iterator hands out
            def __init__(self, v):
                self.v = v

            def numpy(self):
                return self.v

        class PathItr:
            def __init__(self):
                self.it = iter(deal_order())

            def get_next(self):
                try:
                    i = next(self.it)
                except StopIteration:
                    raise tf.errors.OutOfRangeError()
                return tuple(Item(x) for x in paths[i])

        out = {}
        gen = ds._gen_slice_batch(PathItr())
        for n in range(c["n_batches"]):
            cnt, wav, mel, ids = next(gen)
            out["b%d_count" % n] = np.int64(cnt)
            out["b%d_wav" % n] = np.asarray(wav).astype(np.int32)
            out["b%d_mel" % n] = np.asarray(mel).astype(np.float32)
            out["b%d_ids" % n] = np.asarray(ids).astype(np.int32)
    tf._set_eager(False)
    return out


def main():
    tf, tmodel, ops, data = _import_reference()
    for name in TRAIN_CASES:
        out = run_train_case(tf, tmodel, name)
        np.savez_compressed(os.path.join(HERE, "reference_train_%s.npz" % name), **out)
        print(name, "loss", [float(out["s%d_loss" % s]) for s in range(N_STAGES)], "keys", len(out))
    np.savez_compressed(os.path.join(HERE, "reference_mu.npz"), **run_mu(tf, ops))
    np.savez_compressed(os.path.join(HERE, "reference_dealer.npz"), **run_dealer(tf, data))
    print("written")


if __name__ == "__main__":
    main()
