"""GPU tests of the host-side paths around the kernels: the loader's pinned-buffer / copy-stream / event-recycling
path (reference data.py:230-293 surface), a 2-rank NCCL step against the 1-rank step, and
train.py -> checkpoint -> generate.py end to end (reference train.py:216-252, generate.py:32-117)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import wavenet_oracle as O
from tests import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _catalog(n_files, seed, lo=150, hi=900, dtype=np.int32):
    rng = np.random.default_rng(seed)
    return [(int(rng.integers(1, 11)), rng.integers(0, 256, int(rng.integers(lo, hi))).astype(dtype))
            for _ in range(n_files)]


@pytest.mark.parametrize("file_dtype", [np.int32, np.uint8, np.int64])
def test_masked_slice_wav_device_batches_equal_the_dealer(lib, file_dtype):
    """Every batch MaskedSliceWav hands out on the device (uint8 codes + int32 ids copied from pinned buffers on the
    copy stream, codes widened by wn_codes_u8_to_i32) equals, bit for bit, what the slot dealer -- itself bit-exact
    against the oracle's restatement of data.py:110-227 (tests/test_abi_and_host.py) -- deals for the same seed, across
    several recyclings of the 4-buffer ring and with a consumer that lags behind the producer."""
    from lb_wavenet_b200 import data as wdata
    B, T, F = 6, 1000, 120
    cat = _catalog(9, 3, dtype=file_dtype)
    ds = wdata.MaskedSliceWav(None, None, 16000, T, 2, 0, 1, B, 1, "/tmp/t.dset", 0, device="cuda", random_seed=17)
    ds.init_sample_catalog(entries=cat)
    ds.set_receptive_field_size(F)
    ds.build()
    ds.init_vars()
    ref = wdata.SlotDealer(cat, B, T, F, 1, 17, 0, quiet=True)
    junk = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    for n in range(14):
        b = ds.next_batch()
        cnt, w, i = ref.next_batch()
        assert b.wav.dtype == torch.int32 and b.ids.dtype == torch.int32 and b.wav.is_cuda
        if n % 3 == 0:
            junk.zero_()  # keep the compute stream busy: the buffer must not be recycled under the consumer
        got_w, got_i = b.wav.cpu().numpy(), b.ids.cpu().numpy()
        assert b.file_read_count == cnt
        assert np.array_equal(got_w, w), n
        assert np.array_equal(got_i, i), n
    st = ds.loader_stats()
    assert st["h2d_bytes_per_timestep"] == 5.0   # uint8 code + int32 id
    ds._shutdown()


def test_masked_slice_wav_raw_float_input(lib):
    """wav_input_type == 'raw' (reference tmodel.py:59-62): float audio travels as float32 and is mu-law encoded on
    the device by the model; integer transport of such files is refused instead of silently truncating to 0."""
    from lb_wavenet_b200 import data as wdata
    from lb_wavenet_b200._lib import WaveNetLibError
    B, T, F = 3, 400, 50
    rng = np.random.default_rng(5)
    cat = [(int(rng.integers(1, 5)), rng.uniform(-1, 1, int(rng.integers(100, 600))).astype(np.float32)) for _ in range(7)]
    ds = wdata.MaskedSliceWav(None, None, 16000, T, 2, 0, 1, B, 1, "/tmp/t2.dset", 0, device="cuda", random_seed=1,
                              wav_input_type="raw")
    ds.init_sample_catalog(entries=cat)
    ds.set_receptive_field_size(F)
    ds.build()
    ds.init_vars()
    ref = wdata.SlotDealer(cat, B, T, F, 1, 1, 0, quiet=True, wav_dtype="f32")
    for n in range(6):
        b = ds.next_batch()
        _, w, i = ref.next_batch()
        assert b.wav.dtype == torch.float32
        assert np.array_equal(b.wav.cpu().numpy(), w) and np.array_equal(b.ids.cpu().numpy(), i)
    ds._shutdown()
    with pytest.raises(WaveNetLibError):
        wdata.SlotDealer(cat, B, T, F, 1, 1, 0, quiet=True, wav_dtype="u8").next_batch()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


_RANK_SCRIPT = r'''
import os, sys, json, numpy as np, torch
sys.path.insert(0, %(root)r)
from lb_wavenet_b200.dist import DistContext
from lb_wavenet_b200.tmodel import WaveNetTrain, AdamOptimizer
from lb_wavenet_b200.data import SlotDealer
from tests import util
arch = dict(util.CLASSIC_SHALLOW, n_lc_in=0, n_lc_out=0, lc_upsample=[], wav_input_type="mu_law_quant")
B, T = 4, 512
ctx = DistContext.from_env("nccl" if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None)
torch.cuda.set_device(ctx.local_rank)
net = WaveNetTrain(**arch, batch_sz=B, l2_factor=1e-3, add_summary=False, n_keep_checkpoints=1, ckpt_path="/tmp/nccl.net",
                   resume_step=0, n_valid_total=1, print_interval=0, dist=ctx, init_seed=5, device="cuda:%%d" %% ctx.local_rank)
net.build()
net.init_vars()
opt = AdamOptimizer(1e-3)
rng = np.random.default_rng(0)
cat = [(int(rng.integers(1, 11)), rng.integers(0, 256, int(rng.integers(300, 900))).astype(np.int32)) for _ in range(9)]
lo, hi = ctx.slot_range(B)
deal = SlotDealer(cat, B, T, net.get_recep_field_sz(), 1, 5, 0, lo, hi, quiet=True)
losses = []
first = None
for step in range(3):
    _, w, i = deal.next_batch()
    losses.append(net.train_step(torch.as_tensor(w), torch.as_tensor(i), opt))
    if step == 0:
        torch.cuda.synchronize()
        first = {k: v.tolist() for k, v in net.engine.export_state().items() if not k.startswith("SAVE")}
torch.cuda.synchronize()
if ctx.rank == 0:
    st = {k: v.tolist() for k, v in net.engine.export_state().items() if not k.startswith("SAVE")}
    json.dump(dict(losses=losses, state=st, first=first, mode=net._allreduce_mode() if ctx.world > 1 else "none"),
              open(sys.argv[1], "w"))
ctx.barrier()
if ctx.world > 1:
    torch.distributed.destroy_process_group()
'''


@pytest.mark.parametrize("mode", ["single", "buckets"])
def test_two_rank_nccl_step_equals_one_rank_step(lib, tmp_path, mode):
    """3 optimiser steps with the 4 slots sharded over 2 GPUs (NCCL all-reduce of statistics + gradient arena, both
    all-reduce schedules) against the same 3 steps on one GPU.  After the FIRST step -- identical weights going in -- every
    variable agrees to the fp32 atomics' summation-order noise (measured 2e-7 relative: tools/diag_two_rank.py).  Later
    steps are compared as far as they are comparable at all: two runs of the SAME single-GPU program already differ by
    2e-4 (step 2) and 3e-2 (step 3) in the gradients -- a last-bit difference in a weight is re-quantised into whole
    bf16 ulps by the rounding points downstream (profiles/r02e_two_rank_vs_one_rank.md) -- so the
    bound there is the losses (1e-5) and what three Adam steps can move an element."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "rank.py"
    script.write_text(_RANK_SCRIPT % dict(root=ROOT))
    env = dict(os.environ, WN_ALLREDUCE=mode)
    env.pop("RANK", None); env.pop("WORLD_SIZE", None); env.pop("LOCAL_RANK", None)
    one, two = str(tmp_path / "one.json"), str(tmp_path / "two.json")
    subprocess.run([sys.executable, str(script), one], check=True, env=env, cwd=ROOT, timeout=600)
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                    "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script), two],
                   check=True, env=env, cwd=ROOT, timeout=600)
    a, b = json.load(open(one)), json.load(open(two))
    assert b["mode"] == mode
    for la, lb in zip(a["losses"], b["losses"]):
        assert abs(la - lb) <= 1e-5 * abs(la), (a["losses"], b["losses"])
    for k in a["first"]:
        x, y = np.asarray(a["first"][k]), np.asarray(b["first"][k])
        assert util.rel_err(y - 0, x - 0) <= 2e-6, (k, util.rel_err(y - 0, x - 0))
    for k in a["state"]:
        x, y = np.asarray(a["state"][k]), np.asarray(b["state"][k])
        # 3 Adam steps of <= 1e-3 each; an element whose gradient is near zero may take its (sign-like) steps in the
        # other direction (measured: 7e-4 between the 2-rank and a 1-rank run, 3e-4 between two 1-rank runs)
        assert np.abs(x - y).max() <= 2.5e-3, (k, np.abs(x - y).max())


def test_train_cli_checkpoint_then_generate_cli(lib, tmp_path):
    """train.py -> checkpoint -> generate.py with the reference's own command lines (train.py:12-65,
    generate.py:5-29): a catalog TSV of .npy mu-law files, 6 optimiser steps with a save at step 4, then 0.02 s (320 samples)
    for 3 streams written as gen.i{n}.wav (generate.py:114)."""
    rng = np.random.default_rng(0)
    d = tmp_path
    lines = []
    for n in range(6):
        t = np.arange(int(rng.integers(3000, 6000)))
        x = 0.5 * np.sin(t * 0.05 * (n + 1)) + 0.05 * rng.normal(size=t.shape)
        np.save(d / ("f%d.npy" % n), O.mu_encode_np(np.clip(x, -1, 1).astype(np.float32)).astype(np.int32))
        lines.append("%d\t%s\t%s" % (n % 3 + 1, d / ("f%d.npy" % n), "none"))
    (d / "samples.tsv").write_text("\n".join(lines) + "\n")
    arch = dict(n_blocks=2, n_block_layers=5, n_quant=256, n_res=32, n_dil=32, n_skip=256, n_post=256, n_gc_embed=16,
                n_gc_category=3, n_lc_in=0, n_lc_out=0, lc_upsample=[], use_bias=True, wav_input_type="mu_law_quant")
    par = dict(batch_sz=4, sample_rate=16000, slice_sz=1024, l2_factor=1e-3, learning_rate=1e-3, prefetch_sz=2,
               add_summary=False, n_keep_checkpoints=3, n_valid_total=100000)
    (d / "arch.json").write_text(json.dumps(arch))
    (d / "par.json").write_text(json.dumps(par))
    pfx = str(d / "ck")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "train.py"), "-ms", "7", "-si", "4", "-pi", "1", pfx,
                        str(d / "arch.json"), str(d / "par.json"), str(d / "samples.tsv")],
                       capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    assert os.path.exists(pfx + ".net-4.index") and os.path.exists(pfx + ".net-4.data-00000-of-00001"), os.listdir(d)
    rows = [l.split("\t") for l in r.stderr.splitlines() if l.strip() and l.split("\t")[0].strip().isdigit()]
    assert len(rows) >= 5
    losses = [float(x[1]) for x in rows]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]   # it learns
    out = d / "gen"
    out.mkdir()
    r = subprocess.run([sys.executable, os.path.join(ROOT, "generate.py"), "-g", "0.02", "-s", "16000", "-b", "3",
                        "-c", "100", "--seed", "3", str(d / "arch.json"), pfx + ".net-4", str(out)],
                       capture_output=True, text=True, cwd=ROOT, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    import wave
    for n in range(3):
        with wave.open(str(out / ("gen.i%d.wav" % n))) as wf:
            assert wf.getnframes() == 320 and wf.getframerate() == 16000


def test_small_channel_counts_are_zero_padded(lib, tmp_path):
    """reference par/arch2.json has n_res = 3, n_dil = 4, n_skip = 8, n_post = 6: the host mirror zero-extends such
    tensors to what the kernels tile (config.engine_arch) while every checkpoint key keeps the reference's shape.
    The padded model computes exactly the unpadded function: logits and gradients against the oracle on the LOGICAL
    architecture, and the padding is still exactly zero after optimiser steps (L2 + Adam included)."""
    from lb_wavenet_b200 import ckpt
    from lb_wavenet_b200.tmodel import AdamOptimizer, WaveNetTrain
    arch = dict(n_blocks=2, n_block_layers=4, n_quant=256, n_res=3, n_dil=4, n_skip=8, n_post=6, n_gc_embed=16,
                n_gc_category=5, n_lc_in=80, n_lc_out=0, lc_upsample=[4, 4, 4, 4], use_bias=True,
                wav_input_type="mu_law_quant")
    B, T = 3, 256
    net = WaveNetTrain(**arch, batch_sz=B, l2_factor=1e-3, add_summary=False, n_keep_checkpoints=2,
                       ckpt_path=str(tmp_path / "small.net"), resume_step=0, n_valid_total=1000, print_interval=0,
                       init_seed=3)
    net.build()
    net.init_vars()
    eng = net.engine
    assert eng.reg.arch["n_res"] == 32 and eng.reg.arch["n_skip"] == 64
    a = O.Arch(2, 4, 256, 3, 4, 8, 6, 16, 5, True)
    shapes = O.param_shapes(a, B)
    state = {k: net.vars[k].numpy() for k in net.vars}
    for k, (shp, kind) in shapes.items():
        assert tuple(state[k].shape) == tuple(shp), k       # checkpoint keys keep the reference's shapes
    # non-zero biases so that every path is observable
    rng = np.random.default_rng(1)
    for k, (shp, kind) in shapes.items():
        if kind == "bias":
            state[k] = (0.2 * rng.uniform(-1, 1, shp)).astype(np.float32)
            net.vars[k].assign(state[k])
        if kind == "save":
            state[k] = torch.tensor(state[k]).to(torch.bfloat16).float().numpy()
    wav, ids = util.synth_batch(B, T, 5, 9)
    dw, di = torch.as_tensor(wav).cuda(), torch.as_tensor(ids).cuda()
    logits = eng.forward(dw, di, want_logits=True).cpu().numpy()
    eng.backward()
    torch.cuda.synchronize()
    pt, save, kinds = O.to_torch_params(a, state, B, torch.float64, requires_grad=False)
    gem, info = O.train_backward_manual(a, pt, save, torch.as_tensor(wav).long(), torch.as_tensor(ids).long(),
                                        torch.float64, emulate_bf16=True)
    assert np.abs(logits - info["fwd"].logits.numpy()).max() <= 0.05
    assert eng.read_stats()["n_valid"] == info["n_valid"]

    def pad_is_zero(arena):
        for name, pi in eng.reg.params.items():
            full = eng.view(name, arena)
            mask = torch.ones_like(full, dtype=torch.bool)
            mask[tuple(slice(0, d) for d in state[name].shape)] = False
            if bool(mask.any()):
                assert float(full[mask].abs().max()) == 0.0, name

    pad_is_zero(eng.grads)
    for name in ("SIGNAL_0_1", "GATE_1_2", "RESIDUAL_0_0", "SKIP_1_3", "POST1", "POST2", "PRE", "GC_SIGNAL_0_2",
                 "SIGNAL_BIAS_1_0", "POST1_BIAS"):
        g = eng.view(name, eng.grads)[tuple(slice(0, d) for d in state[name].shape)].cpu().numpy()
        assert util.rel_err(g, gem[name].numpy()) <= 6e-2, (name, util.rel_err(g, gem[name].numpy()))
    opt = AdamOptimizer(1e-3)
    for _ in range(3):
        net.train_step(dw, di, opt)
    pad_is_zero(eng.params)
    pad_is_zero(eng.m)
    path = net.save(3)
    tensors = ckpt.read_checkpoint(path)
    for k, (shp, kind) in shapes.items():
        assert tuple(tensors[k].shape) == tuple(shp), k
    assert tuple(tensors["SIGNAL_0_1/Adam"].shape) == (2, 3, 4)
