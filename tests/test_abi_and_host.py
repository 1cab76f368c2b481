"""CPU tests: the C-ABI library loads and exports what include/*.h declares, the variable registry
matches the reference's names/shapes, and the host-side mirror (config, checkpoint, loader, CLIs)
behaves like the reference's.  No compute call is made (no GPU here)."""
import ctypes as C
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import wavenet_oracle as O
from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the reference's shipped arch/par files, restated as dicts (par/arch1-5.json, par/par1-3.json)
REF_ARCH = {
    "arch1": dict(n_blocks=5, n_block_layers=10, n_quant=256, n_res=32, n_dil=32, n_skip=512, n_post1=512,
                  n_gc_embed=17, n_gc_category=377),
    "arch2": dict(n_blocks=5, n_block_layers=10, n_quant=256, n_res=3, n_dil=4, n_skip=8, n_post=6, n_gc_embed=16,
                  n_lc_in=80, n_lc_out=0, lc_upsample=[4, 4, 4, 4], use_bias=True, wav_input_type="mu_law_quant"),
    "arch3": dict(n_blocks=5, n_block_layers=10, n_quant=256, n_res=32, n_dil=32, n_skip=512, n_post1=512,
                  n_gc_embed=0, n_gc_category=0, use_bias=True),
    "arch4": dict(n_blocks=5, n_block_layers=10, n_quant=256, n_res=32, n_dil=32, n_skip=512, n_post=512,
                  n_gc_embed=16, n_lc_in=80, n_lc_out=80, lc_hop_sz=256, lc_upsample=[4, 4, 4, 4], use_bias=True,
                  wav_input_type="mu_law_quant"),
    "arch5": dict(n_blocks=5, n_block_layers=10, n_quant=256, n_res=32, n_dil=32, n_skip=512, n_post=512,
                  n_gc_embed=16, n_gc_category=376, n_lc_in=80, n_lc_out=80, lc_upsample=[4, 4, 4, 4], use_bias=True,
                  wav_input_type="mu_law_quant"),
}
REF_PAR = {
    "par1": dict(batch_sz=10, sample_rate=16000, slice_sz=512, l2_factor=0.001, learning_rate=0.001, prefetch_sz=10,
                 add_summary=False, n_keep_checkpoints=10, n_valid_total=3263771401),
    "par3": dict(batch_sz=1, sample_rate=16000, slice_sz=1024, l2_factor=0.001, learning_rate=0.001, prefetch_sz=10,
                 max_to_keep=30),
}


# ---------------------------------------------------------------- C ABI ------------------------------
def test_library_exports_every_declared_symbol(lib):
    from lb_wavenet_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "wavenet_b200.h")).read()
    declared = set(re.findall(r"\b(wn_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), "symbol %s declared in the header but not exported" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.wn_abi_version() == 2
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (wn_[a-z0-9_]+)", out))
    assert declared <= exported


@pytest.mark.parametrize("arch", [util.TINY, util.TINY_GC, util.TINY_NOBIAS, util.WIDE, util.CLASSIC, util.C1,
                                  util.TINY_LC, util.TINY_GC_LC, util.ARCH5])
def test_registry_matches_reference_names_and_shapes(lib, arch):
    """arch.py:85-103,126,142 + tmodel.py:292-328: serial names, shapes, construction order."""
    from lb_wavenet_b200.engine import Registry
    B = 3
    reg = Registry(arch, B)
    a = util.oracle_arch(arch)
    ref = O.param_shapes(a, B)
    train = [(k, s) for k, (s, kind) in ref.items() if kind in ("filter", "bias")]
    assert [(n, i.shape) for n, i in reg.params.items()] == train
    for (k, (s, kind)), info in zip([kv for kv in ref.items() if kv[1][1] in ("filter", "bias")], reg.params.values()):
        assert info.kind == (0 if kind == "filter" else 1)
    saves = [(k, s) for k, (s, kind) in ref.items() if kind == "save"]
    assert [(s.name, s.shape) for s in reg.saves] == saves
    assert reg.recep_field == a.recep_field() and reg.n_layers == a.n_layers
    # arena: non-overlapping, aligned
    spans = sorted((i.offset, i.offset + i.numel) for i in reg.params.values())
    assert all(b0 >= a1 for (_, a1), (b0, _) in zip(spans, spans[1:])) and spans[-1][1] <= reg.n_param_elems
    assert all(o % 64 == 0 for o, _ in spans)
    assert reg.save_elems == sum(int(np.prod(s)) for _, s in saves)
    hop = reg.lc_hop  # with local conditioning the stage length is a multiple of prod(lc_upsample) (data.py:32-37)
    assert reg.workspace_bytes(64 * hop) < reg.workspace_bytes(128 * hop)


def test_unsupported_architectures_are_rejected_with_a_message(lib):
    from lb_wavenet_b200 import _lib
    from lb_wavenet_b200.engine import Registry
    for bad in (dict(util.TINY, n_quant=128), dict(util.TINY, n_res=3), dict(util.TINY, n_skip=1024),
                dict(util.TINY, n_gc_embed=4, n_gc_category=0)):
        with pytest.raises(_lib.WaveNetLibError) as e:
            Registry(bad, 1)
        assert "unsupported architecture" in str(e.value)
    with pytest.raises(_lib.WaveNetLibError):
        Registry(util.TINY, 0)


def test_no_cpu_fallback_and_no_oracle_in_the_product_path(lib):
    import torch
    from lb_wavenet_b200 import _lib
    from lb_wavenet_b200.engine import TrainEngine
    if not torch.cuda.is_available():
        with pytest.raises(_lib.WaveNetLibError):
            TrainEngine(util.TINY, 1)
    pkg = os.path.join(ROOT, "lb_wavenet_b200")
    for fn in [os.path.join(pkg, f) for f in os.listdir(pkg) if f.endswith(".py")] + \
            [os.path.join(ROOT, f) for f in ("train.py", "generate.py", "slice_data.py")]:
        src = open(fn).read()
        assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), fn


# ---------------------------------------------------------------- config -----------------------------
def test_arch_schema_drift_is_normalised():
    from lb_wavenet_b200 import config
    a1 = config.normalize_arch(REF_ARCH["arch1"], warn=False)
    assert a1["n_post"] == 512 and "n_post1" not in a1 and a1["use_bias"] is True
    assert a1["n_lc_out"] == 0 and a1["lc_upsample"] == [] and a1["wav_input_type"] == "mu_law_quant"
    assert config.mel_hop_sz(a1) == 1
    with pytest.raises(config.ConfigError):
        config.normalize_arch(REF_ARCH["arch2"], warn=False)  # needs --num-global-cond (train.py:81-84)
    a2 = config.normalize_arch(REF_ARCH["arch2"], num_global_cond=12, warn=False)
    assert a2["n_gc_category"] == 12 and config.mel_hop_sz(a2) == 256
    a4 = config.normalize_arch(REF_ARCH["arch4"], num_global_cond=5, warn=False)
    assert "lc_hop_sz" not in a4
    e4 = config.engine_arch(a4)  # local conditioning: the upsampling strides and channel counts reach the C ABI
    assert e4["n_lc_in"] == 80 and e4["n_lc_out"] == 80 and e4["lc_upsample"] == [4, 4, 4, 4]
    e2 = config.engine_arch(a2)  # arch2: n_lc_out == 0 -> no LC; tiny channel counts are zero-padded
    assert "n_lc_out" not in e2 and (e2["n_res"], e2["n_dil"], e2["n_skip"], e2["n_post"]) == (32, 32, 64, 64)
    a3 = config.normalize_arch(REF_ARCH["arch3"], warn=False)
    # the normalised dict is exactly what WaveNetTrain.__init__ consumes (tmodel.py:8-24)
    assert set(a3) == {"n_blocks", "n_block_layers", "n_quant", "n_res", "n_dil", "n_skip", "n_post", "n_gc_embed",
                       "n_gc_category", "n_lc_in", "n_lc_out", "lc_upsample", "use_bias", "wav_input_type"}
    p3 = config.normalize_par(REF_PAR["par3"])
    assert p3["n_keep_checkpoints"] == 30 and p3["add_summary"] is False and "max_to_keep" not in p3
    assert config.normalize_par(REF_PAR["par1"])["n_valid_total"] == 3263771401
    for f in os.listdir(os.path.join(ROOT, "par")):
        path = os.path.join(ROOT, "par", f)
        (config.load_arch if f.startswith("arch") else config.load_par)(path)


# ---------------------------------------------------------------- checkpoint -------------------------
def test_checkpoint_naming_roundtrip_and_pruning(tmp_path):
    """ckpt.py:8-11,41-42,54-62,70-76."""
    from lb_wavenet_b200 import ckpt
    store = {"PRE": np.arange(6, dtype=np.float32).reshape(2, 3), "GLOBAL_STEP": np.array(7, np.int32),
             "SAVE_4_0_2": np.ones((2, 4, 3), np.float32)}

    def var(k):
        return ckpt.Variable(k, store[k].shape, store[k].dtype, lambda: store[k],
                             lambda v: store.__setitem__(k, v.copy()))

    c = ckpt.Checkpoint(str(tmp_path / "run.net"), 2, 0)
    with pytest.raises(ValueError):
        c.save(1)
    c.add_saveable_objects({k: var(k) for k in store})
    for step in (10, 20, 30):
        store["GLOBAL_STEP"] = np.array(step, np.int32)
        pfx = c.save(step)
        assert pfx == str(tmp_path / "run.net") + "-%d" % step
        for s in ("index", "meta", "data-00000-of-00001"):
            assert os.path.exists(pfx + "." + s)
    assert not os.path.exists(str(tmp_path / "run.net-10.index"))  # max_to_keep = 2
    assert 'model_checkpoint_path: "run.net-30"' in open(tmp_path / "checkpoint").read()
    # the data file is the raw little-endian tensors in key order (TF bundle payload layout)
    raw = open(str(tmp_path / "run.net-30.data-00000-of-00001"), "rb").read()
    assert raw[:4] == np.array(30, "<i4").tobytes()
    store["PRE"] = np.zeros((2, 3), np.float32)
    c.resume_step = 20
    c.restore()
    assert store["PRE"][1, 2] == 5 and int(store["GLOBAL_STEP"]) == 20
    c.resume_step = 10
    with pytest.raises(SystemExit):
        c.restore()


# ---------------------------------------------------------------- loader -----------------------------
def _catalog(rng, n, lo, hi):
    return [(int(rng.integers(1, 9)), rng.integers(0, 256, int(rng.integers(lo, hi))).astype(np.int32))
            for _ in range(n)]


def _ref_batches(cat, B, T, F, hop, seed, pos):
    def files():
        cnt = pos
        for idx in O.shuffled_repeat_order(len(cat), seed, pos):
            cnt += 1
            yield cnt, cat[idx][0], cat[idx][1]
    return O.gen_slice_batches(files(), B, T, F, hop)


@pytest.mark.parametrize("B,T,F,hop,lo,hi", [(5, 64, 31, 4, 5, 400), (1, 33, 10, 1, 10, 40), (8, 512, 46, 1, 2000, 9000),
                                             (3, 16, 16, 16, 16, 64), (4, 100, 7, 1, 7, 8)])
def test_slot_dealer_bit_exact_vs_reference_generators(B, T, F, hop, lo, hi):
    """data.py:110-227: windows, id masks, file->slot dealing, read counts -- ragged lengths, files shorter
    than F (skipped), hop trimming, windows spanning several files, files spanning several windows."""
    from lb_wavenet_b200.data import SlotDealer
    rng = np.random.default_rng(B * 1000 + T)
    cat = _catalog(rng, 19, lo, hi)
    ref = _ref_batches(cat, B, T, F, hop, 7, 3)
    whole = SlotDealer(cat, B, T, F, hop, 7, 3, quiet=True)
    shards = [SlotDealer(cat, B, T, F, hop, 7, 3, slot_lo=s, slot_hi=min(B, s + 2), quiet=True) for s in range(0, B, 2)]
    for it in range(60):
        rc, rw, ri = next(ref)
        c, w, i = whole.next_batch()
        assert c == rc and np.array_equal(w, rw) and np.array_equal(i, ri), it
        assert w.dtype == np.int32 and i.dtype == np.int32  # data.py:262-265
        for k, sh in enumerate(shards):  # data-parallel replay: each rank materialises only its slots
            c2, w2, i2 = sh.next_batch()
            assert c2 == rc and np.array_equal(w2, rw[2 * k:2 * k + 2]) and np.array_equal(i2, ri[2 * k:2 * k + 2])


def test_loader_surface_resume_and_alignment(tmp_path):
    """MaskedSliceWav: slice_sz rounded UP to the hop (data.py:32-37), TSV catalog (data.py:43-48), save /
    restore of (random_seed, ckpt_position) (data.py:273-286), background producer == direct dealing."""
    from lb_wavenet_b200 import data
    rng = np.random.default_rng(1)
    rows = []
    for k in range(6):
        wav = rng.integers(0, 256, int(rng.integers(300, 900))).astype(np.int32)
        wp, mp = tmp_path / ("f%d.wav.npy" % k), tmp_path / ("f%d.mel.npy" % k)
        np.save(wp, wav)
        np.save(mp, np.zeros((len(wav) // 4, 2), np.float32))
        rows.append("%d\t%s\t%s" % (k + 1, wp, mp))
    (tmp_path / "cat.tsv").write_text("\n".join(rows) + "\n")
    ds = data.MaskedSliceWav(None, str(tmp_path / "cat.tsv"), 16000, 130, 2, 2, 4, 3, 5, str(tmp_path / "r.dset"), 0,
                             device="cpu", random_seed=5)
    assert ds.slice_sz == 132
    ds.init_sample_catalog()
    assert ds.get_max_id() == 6
    with pytest.raises(ValueError):
        ds.build()
    ds.set_receptive_field_size(30)
    ds.build()
    ds.init_vars()
    cat = [(r[0], np.load(r[1])) for r in ds.sample_catalog]
    ref = _ref_batches(cat, 3, 132, 30, 4, 5, 0)
    for _ in range(7):
        b = ds.next_batch()
        rc, rw, ri = next(ref)
        assert b.file_read_count == rc and np.array_equal(b.wav, rw) and np.array_equal(b.ids, ri)
    cnt, wav, mel, ids = next(ds.get_itr())
    next(ref)
    pfx = ds.save(40, cnt)
    assert pfx.endswith("r.dset-40") and ds.ckpt_position == cnt
    ds2 = data.MaskedSliceWav(None, str(tmp_path / "cat.tsv"), 16000, 132, 2, 2, 4, 3, 5, str(tmp_path / "r.dset"), 40,
                              device="cpu")
    ds2.init_sample_catalog()
    ds2.set_receptive_field_size(30)
    ds2.build()
    ds2.init_vars()
    ds2.restore()
    assert ds2.random_seed == 5 and ds2.ckpt_position == cnt
    # our own checkpoints carry the dealer's exact state as optional extra keys: the resumed
    # stream continues exactly where the consumed one stopped -- every slot mid-file, same file-to-slot dealing
    assert sorted(ds2.restored_optional) == ["slot_count", "slot_file", "slot_pos", "stream_position"]
    for _ in range(5):
        b = ds2.next_batch()
        rc, rw, ri = next(ref)
        assert b.file_read_count == rc and np.array_equal(b.wav, rw) and np.array_equal(b.ids, ri)
    # a checkpoint with only the reference's two scalars (data.py:273-276): the reference's approximate resume, i.e.
    # the file stream restarts at skip(ckpt_position) and every slot starts a fresh file (data.py:249-250)
    from lb_wavenet_b200 import tfbundle
    tfbundle.write_bundle(str(tmp_path / "r.dset-41"), {"random_seed": np.array(5, np.int64),
                                                         "ckpt_position": np.array(cnt, np.int64)})
    (tmp_path / "r.dset-41.meta").write_bytes(b"")
    ds3 = data.MaskedSliceWav(None, str(tmp_path / "cat.tsv"), 16000, 132, 2, 2, 4, 3, 5, str(tmp_path / "r.dset"), 41,
                              device="cpu")
    ds3.init_sample_catalog()
    ds3.set_receptive_field_size(30)
    ds3.build()
    ds3.restore()      # before init_vars(), as train.py may do: the start-up honours the restored position
    ds3.init_vars()
    assert ds3.restored_optional == []
    ref2 = _ref_batches(cat, 3, 132, 30, 4, 5, cnt)
    b = ds3.next_batch()
    rc, rw, ri = next(ref2)
    assert b.file_read_count == rc and np.array_equal(b.wav, rw)
    for d in (ds, ds2, ds3):
        d._shutdown()


def test_checkpoint_optional_keys(tmp_path):
    """Optional variables are written, restored when present and skipped (not an error) when a checkpoint lacks them;
    a missing mandatory key is still an error (reference ckpt.py:65-81 restores exactly its own dict)."""
    from lb_wavenet_b200 import ckpt
    store = {"W": np.arange(6, dtype=np.float32).reshape(2, 3), "W/Adam": np.ones((2, 3), np.float32)}

    def var(name, optional):
        return ckpt.Variable(name, (2, 3), np.float32, lambda: store[name], lambda v: store.__setitem__(name, v.copy()),
                             optional=optional)
    c = ckpt.Checkpoint(str(tmp_path / "a.net"), 2, 7)
    c.add_saveable_objects({"W": var("W", False)})
    c.save(7)                                    # a "reference" checkpoint: no slot keys
    c.add_saveable_objects({"W/Adam": var("W/Adam", True)})
    store["W"] = np.zeros((2, 3), np.float32)
    c.restore()
    assert c.restored_optional == [] and store["W"][1, 2] == 5 and store["W/Adam"][0, 0] == 1
    store["W/Adam"] = np.full((2, 3), 3, np.float32)
    c.save(8)
    store["W/Adam"] = np.zeros((2, 3), np.float32)
    c.restore(str(tmp_path / "a.net-8"))
    assert c.restored_optional == ["W/Adam"] and store["W/Adam"][1, 1] == 3
    c2 = ckpt.Checkpoint(str(tmp_path / "a.net"), 2, 8)
    c2.add_saveable_objects({"W": var("W", False), "V": var("W", False)})
    with pytest.raises(KeyError):
        c2.restore()


def test_slice_data_cli(tmp_path):
    """slice_data.py:6-24,47-99: hop-aligned cut, too-short files skipped, new catalog written."""
    import slice_data
    rows = []
    for k, n in enumerate((4096, 1000)):
        wp, mp = tmp_path / ("s%d.wav.npy" % k), tmp_path / ("s%d.mel.npy" % k)
        np.save(wp, np.arange(n, dtype=np.int32))
        np.save(mp, np.arange(n // 256, dtype=np.float32)[:, None])
        rows.append("%d\t%s\t%s" % (k + 1, wp, mp))
    if 1000 % 256:
        np.save(tmp_path / "s1.mel.npy", np.zeros((1000 // 256, 1), np.float32))
    (tmp_path / "in.rdb").write_text("\n".join(rows) + "\n")
    out_dir = tmp_path / "out"
    slice_data.main(["-hs", "256", "-sp", "300", "-ss", "1100", str(tmp_path / "in.rdb"), str(out_dir),
                     str(tmp_path / "out.rdb")])
    lines = (tmp_path / "out.rdb").read_text().strip().split("\n")
    assert len(lines) == 1 and lines[0].startswith("1\t")
    w = np.load(lines[0].split("\t")[1])
    m = np.load(lines[0].split("\t")[2])
    assert w[0] == 256 and len(w) == 1024 and len(m) == 4 and m[0, 0] == 1.0  # beg 300->256, size 1100->1024


def test_wav_io_roundtrip(tmp_path):
    from lb_wavenet_b200 import wavio
    x = (0.5 * np.sin(np.arange(3200) * 0.05)).astype(np.float32)
    wavio.write_wav(str(tmp_path / "a.wav"), x, 16000)
    y = wavio.read_wav(str(tmp_path / "a.wav"), 16000, 0.05, 0.1)
    assert y.shape == (1600,) and np.abs(y - x[800:2400]).max() < 1e-4
    with pytest.raises(ValueError):
        wavio.read_wav(str(tmp_path / "a.wav"), 22050)


def test_cli_surfaces_accept_the_reference_flags():
    """train.py:12-65, generate.py:5-29."""
    import generate
    import train
    a = train.get_args(["-tf", "t.json", "-pd", "d", "-rs", "3", "-s", "-cpu", "-tb", "d", "-si", "5", "-pi", "2",
                        "-tdb", "-ms", "9", "-bs", "4", "-ss", "64", "-l2", "0.1", "-lr", "0.01", "-gc", "7",
                        "pfx", "a.json", "p.json", "s.tsv"])
    assert (a.resume_step, a.save_interval, a.max_steps, a.batch_size, a.num_global_cond) == (3, 5, 9, 4, 7)
    assert train.get_args(["pfx", "a", "p", "s"]).save_interval == 1000
    g = generate.get_args(["-w", "t.wav", "-ts", "0.5", "-td", "1", "-g", "2", "-s", "8000", "-c", "100", "-b", "3",
                           "a.json", "ck", "out"])
    assert (g.gen_seconds, g.sample_rate, g.chunk_size, g.batch_size) == (2.0, 8000, 100, 3)
    d = generate.get_args(["a.json", "ck", "out"])
    assert (d.gen_seconds, d.sample_rate, d.chunk_size, d.batch_size) == (5, 16000, 1000, 10)


def test_bucket_plan_covers_the_arena_once(lib):
    from lb_wavenet_b200.dist import bucket_plan
    from lb_wavenet_b200.engine import Registry
    for arch in (util.CLASSIC, util.C1, util.TINY):
        reg = Registry(arch, 2)
        plan = bucket_plan([(n, i.offset, i.numel) for n, i in reg.params.items()], reg.n_layers,
                           arch["n_block_layers"], reg.n_param_elems, arch["n_gc_embed"] > 0)
        cover = np.zeros(reg.n_param_elems, np.int32)
        last = 0
        for phase_end, ranges in plan:
            assert phase_end > last
            last = phase_end
            for lo, hi in ranges:
                cover[lo:hi] += 1
        assert last == reg.n_layers + 2 and (cover == 1).all()
        # a layer's gradients are only reduced after the phase that finalises them
        nbl = arch["n_block_layers"]
        for name, info in reg.params.items():
            parts = name.split("_")
            if parts[-1].isdigit() and parts[-2].isdigit():
                layer = int(parts[-2]) * nbl + int(parts[-1])
                need = reg.n_layers - layer + 1 if arch["n_gc_embed"] == 0 else reg.n_layers + 2
                pe = next(pe for pe, rs in plan if any(lo <= info.offset < hi for lo, hi in rs))
                assert pe >= need, (name, pe, need)
