"""The oracle -- and, on a GPU, the CUDA path -- against vectors produced by the REFERENCE'S OWN PYTHON.

tests/golden/reference_*.npz were written by tests/golden/make_reference_vectors.py, which imports tmodel.py / arch.py /
ops.py / data.py unmodified from the reference tree and runs them on oracle/tf1_shim (TensorFlow itself is not
installable here).  Inputs and parameters are rebuilt from the seeds that script exports; nothing here reads
/root/reference.
"""
import importlib.util
import os

import numpy as np
import pytest
import torch

from oracle import wavenet_oracle as O
from tests import util

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_reference_vectors", os.path.join(HERE, "golden", "make_reference_vectors.py"))
G = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(G)


def _golden(name):
    return np.load(os.path.join(HERE, "golden", "reference_%s.npz" % name))


def _close(a, b, rtol, atol=0.0):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() <= atol + rtol * max(np.abs(b).max(), 1e-300) if a.size else True


# ---------------------------------------------------------------- training graph (tmodel.py:284-339) -------------
@pytest.mark.parametrize("name", sorted(G.TRAIN_CASES))
@pytest.mark.parametrize("conv_impl", ["taps", "conv1d"])
def test_oracle_training_graph_equals_the_reference_run(name, conv_impl):
    """Two consecutive stages with the SAVE state carried by the reference itself: loss, logits, every gradient (a
    strided sample of each tensor + its norm), the SAVE variables after each stage and the two counters.  fp64 on both
    sides: agreement to 1e-9."""
    arch, B, T, l2, _ = G.TRAIN_CASES[name]
    a = G.oracle_arch(arch)
    gold = _golden("train_" + name)
    p = G.train_params(name)
    # variable names and creation order are the checkpoint contract (arch.py:112-142)
    assert list(gold["var_order"]) == [k for k in O.param_shapes(a, B).keys()]
    n_valid_cumul = 0
    for stage in range(G.N_STAGES):
        wav, ids, mel = G.train_inputs(name, stage)
        pt, save, kinds = O.to_torch_params(a, p, B, torch.float64)
        w, i = torch.as_tensor(wav).long(), torch.as_tensor(ids).long()
        mt = None if mel is None else torch.as_tensor(mel, dtype=torch.float64)
        fwd = O.train_forward(a, pt, save, w, i, torch.float64, conv_impl=conv_impl, mel=mt)
        L = O.loss_fn(a, fwd.logits, w, i, pt, kinds, l2)
        L.total.backward()
        s = "s%d_" % stage
        assert abs(float(L.total) - float(gold[s + "loss"])) <= 1e-10 * abs(float(gold[s + "loss"])), (stage, float(L.total))
        assert _close(fwd.logits.detach().numpy()[:, ::G.logit_stride(name), :], gold[s + "logits"], 1e-6)  # stored as float32
        for k, v in pt.items():
            g = v.grad.numpy() if v.grad is not None else np.zeros(tuple(v.shape))
            ref = gold[s + "grad_" + k]
            assert _close(G.sample_of(g, G.sample_limit(name)), ref, 1e-9, 1e-14), (stage, k)
            assert abs(np.sqrt((g ** 2).sum()) - float(gold[s + "gradnorm_" + k])) <= 1e-9 * max(float(gold[s + "gradnorm_" + k]), 1e-12), (stage, k)
        n_valid_cumul += L.n_valid
        assert int(gold[s + "global_step"]) == stage + 1  # tmodel.py:276
        assert int(gold[s + "valid_samples"]) == n_valid_cumul  # tmodel.py:277
        for li, ((b, bl), dil) in enumerate(zip(a.layer_ids(), a.dilations())):
            key = "SAVE_%d_%d_%d" % (dil, b, bl)
            assert _close(G.save_sample(fwd.new_save[li].numpy()), gold[s + key], 1e-12, 1e-15), (stage, key)
            p[key] = fwd.new_save[li].numpy()  # carried into the next stage (tmodel.py:165)


def test_reference_run_covers_the_quirks_the_oracle_documents():
    """The r32 case holds an out-of-range code at a valid position: in the reference run its label row is all zero, so
    it adds nothing to the loss yet is counted in n_valid, and its logits row still receives softmax as gradient (TF's
    fused xent kernel) -- the oracle statement of that (loss_fn) is what the test above compares gradients through.
    Here: dropping the position from the mask changes the oracle's loss, i.e. the case is sensitive to the quirk."""
    name = "r32"
    arch, B, T, l2, _ = G.TRAIN_CASES[name]
    a = G.oracle_arch(arch)
    p = G.train_params(name)
    wav, ids, _ = G.train_inputs(name, 0)
    assert wav[0, 9] == -1 and ids[0, 9] != 0
    pt, save, kinds = O.to_torch_params(a, p, B, torch.float64, requires_grad=False)
    w = torch.as_tensor(wav).long()
    fwd = O.train_forward(a, pt, save, w, torch.as_tensor(ids).long(), torch.float64)
    L1 = O.loss_fn(a, fwd.logits, w, torch.as_tensor(ids).long(), pt, kinds, l2)
    ids2 = ids.copy()
    ids2[0, 9] = 0
    L2 = O.loss_fn(a, fwd.logits, w, torch.as_tensor(ids2).long(), pt, kinds, l2)
    assert L1.n_valid == L2.n_valid + 1 and abs(float(L1.total) - float(L2.total)) > 1e-4
    assert abs(float(L1.total) - float(_golden("train_r32")["s0_loss"])) < 1e-9


@pytest.mark.parametrize("name", sorted(G.TRAIN_CASES))
def test_product_registry_equals_the_reference_runs_variable_list(name):
    """wn_param_info / wn_save_info (the C ABI's registry; no GPU needed) against `net.vars` of the reference run: the same
    serial names in the same creation order -- the checkpoint contract (arch.py:112-142, tmodel.py:292-328)."""
    from lb_wavenet_b200 import config
    from lb_wavenet_b200.engine import Registry
    arch, B, T, _, _ = G.TRAIN_CASES[name]
    order = [str(k) for k in _golden("train_" + name)["var_order"]]
    reg = Registry(config.engine_arch(dict(arch)), B)  # (tiny channel counts are zero-padded: names and order are what is compared)
    ours, it_p, it_s = [], iter(reg.params), iter(reg.saves)
    for k in order:  # trainable variables and SAVE state are two lists in the ABI: merge them along the reference's order
        if k in ("GLOBAL_STEP", "VALID_SAMPLES"):
            continue  # host-side counters of the mirror (tmodel.py:223-226)
        ours.append(next(it_s).name if k.startswith("SAVE") else next(it_p))
    assert ours == [k for k in order if k not in ("GLOBAL_STEP", "VALID_SAMPLES")]
    assert next(it_p, None) is None and next(it_s, None) is None


def test_the_vectors_have_teeth():
    """Mutations of the kind a misreading of the reference would produce move the loss far outside the 1e-10 the comparison
    allows: the two conv taps swapped in one layer, the SAVE rows of one layer rotated by one timestep, another voice id,
    one input code changed, one bias scaled by 0.1 %."""
    name = "gc"
    arch, B, T, l2, _ = G.TRAIN_CASES[name]
    a = G.oracle_arch(arch)
    gold = float(_golden("train_" + name)["s0_loss"])
    p = G.train_params(name)
    wav, ids, _ = G.train_inputs(name, 0)

    def loss(p_, wav_, ids_):
        pt, save, kinds = O.to_torch_params(a, p_, B, torch.float64, requires_grad=False)
        w, i = torch.as_tensor(wav_).long(), torch.as_tensor(ids_).long()
        return float(O.loss_fn(a, O.train_forward(a, pt, save, w, i, torch.float64).logits, w, i, pt, kinds, l2).total)
    assert abs(loss(p, wav, ids) - gold) <= 1e-10 * gold
    ids2 = ids.copy()
    ids2[ids2 == 2] = 3
    wav2 = wav.copy()
    wav2[1, 5] = (wav2[1, 5] + 1) % 256
    mutants = {
        "taps swapped": (dict(p, SIGNAL_0_1=p["SIGNAL_0_1"][::-1].copy()), wav, ids),
        "SAVE rotated": (dict(p, SAVE_2_0_1=np.roll(p["SAVE_2_0_1"], 1, axis=1)), wav, ids),
        "voice id": (p, wav, ids2),
        "input code": (p, wav2, ids),
        "bias scaled": (dict(p, POST1_BIAS=p["POST1_BIAS"] * 1.001), wav, ids),
    }
    for what, (p_, w_, i_) in mutants.items():
        assert abs(loss(p_, w_, i_) - gold) > 1e-6, what


# ---------------------------------------------------------------- mu-law (ops.py:4-39) ---------------------------
def test_mu_law_equals_the_reference_run():
    gold = _golden("mu")
    x = G.mu_inputs()
    # the reference's two statements: the numpy twins (ops.py:23-39) run as they are, the tf ones through the shim in float32
    assert np.array_equal(O.mu_encode_np(x), gold["codes_np"]) or _mu_diff_is_numpy2_promotion(x, gold["codes_np"])
    assert np.array_equal(O.mu_encode_np(x), gold["codes_tf"])
    dec = O.mu_decode_np(np.arange(256))
    assert np.array_equal(dec, gold["decoded_np"])  # ops.mu_decode_np run as it is
    # ops.mu_decode through the shim: torch's float32 pow against numpy's, a last-bit difference of the shim's arithmetic
    assert np.abs(dec - gold["decoded_tf"]).max() <= 1.2e-7


def _mu_diff_is_numpy2_promotion(x, codes_np):
    """ops.mu_encode_np run under numpy >= 2 divides by the float64 scalar log1p(mu) in float64 (NEP 50 keeps float32
    only for python scalars) -- the reference was written against numpy 1.x, where every intermediate stays float32
    (oracle mu_encode_np docstring).  Differences, if any, are single codes at float32 threshold neighbours."""
    d = O.mu_encode_np(x).astype(np.int64) - codes_np.astype(np.int64)
    return np.abs(d).max() <= 1 and (d != 0).mean() < 1e-3


# ---------------------------------------------------------------- slot dealer (data.py:110-227) ------------------
def _deal_catalog():
    return G.deal_files()


def test_oracle_dealer_equals_the_reference_run():
    c = G.DEAL_CASE
    gold = _golden("dealer")
    files, order = _deal_catalog(), G.deal_order()

    def file_iter():
        for cnt, idx in enumerate(order, start=1):
            yield cnt, files[idx][0], files[idx][1]
    T = O.align_slice_sz(c["slice_sz"], c["mel_hop_sz"])
    gen = O.gen_slice_batches(file_iter(), c["batch_sz"], T, c["recep_field_sz"], c["mel_hop_sz"])
    for n in range(c["n_batches"]):
        cnt, wav, ids = next(gen)
        assert cnt == int(gold["b%d_count" % n]), n
        assert np.array_equal(wav, gold["b%d_wav" % n]) and np.array_equal(ids, gold["b%d_ids" % n]), n


def test_product_dealer_equals_the_reference_run():
    """the C slot dealer behind MaskedSliceWav (wn_deal_plan / wn_deal_fill), mel frames included, whole and sharded"""
    from lb_wavenet_b200.data import SlotDealer
    c = G.DEAL_CASE
    gold = _golden("dealer")
    files, order = _deal_catalog(), G.deal_order()
    B, T = c["batch_sz"], O.align_slice_sz(c["slice_sz"], c["mel_hop_sz"])

    def dealer(lo, hi):
        d = SlotDealer(files, B, T, c["recep_field_sz"], c["mel_hop_sz"], 0, 0, slot_lo=lo, slot_hi=hi, quiet=True,
                       mel_channels=c["mel_spectrum_sz"])
        d._order = iter(int(i) for i in order)  # the reference run's file order instead of the dealer's own shuffle
        d._order_buf = np.empty(0, np.int32)
        return d
    whole, shard = dealer(0, B), dealer(1, 3)
    for n in range(c["n_batches"]):
        cnt, wav, ids = whole.next_batch()
        assert cnt == int(gold["b%d_count" % n]), n
        assert np.array_equal(wav, gold["b%d_wav" % n]) and np.array_equal(ids, gold["b%d_ids" % n]), n
        assert np.array_equal(whole.last_mel, gold["b%d_mel" % n]), n
        cnt2, wav2, ids2 = shard.next_batch()
        assert cnt2 == cnt and np.array_equal(wav2, wav[1:3]) and np.array_equal(ids2, ids[1:3])
        assert np.array_equal(shard.last_mel, gold["b%d_mel" % n][1:3])


# ---------------------------------------------------------------- CUDA path --------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(G.TRAIN_CASES))
def test_cuda_training_step_against_the_reference_run(name, tmp_path):
    """wn_train_forward / wn_train_backward behind the host mirror of the reference's WaveNetTrain, on the reference
    run's inputs, two stages with carried SAVE: logits and loss within the bf16 contract of DESIGN.md section 4 (operands
    and stored activations are bf16, the reference is fp32 -- here fp64), every gradient tensor within the bound the
    emulated-oracle tests establish, SAVE after each stage.  ('odd' runs zero-padded: config.engine_arch.)"""
    from lb_wavenet_b200.tmodel import WaveNetTrain
    arch, B, T, l2, _ = G.TRAIN_CASES[name]
    gold = _golden("train_" + name)
    p = G.train_params(name)
    raw = name in G.RAW_CASES  # float audio in, mu-law encoded on the device (wn_mu_encode) as tmodel.py:59-62 does in the graph
    net = WaveNetTrain(**arch, wav_input_type="raw" if raw else "mu_law_quant", batch_sz=B, l2_factor=l2, add_summary=False,
                       n_keep_checkpoints=1, ckpt_path=str(tmp_path / "ref.net"), resume_step=0, n_valid_total=10 ** 6,
                       print_interval=0, init_seed=1)
    net.build()
    net.init_vars()
    eng = net.engine
    a = G.oracle_arch(arch)
    shapes = O.param_shapes(a, B)
    for k, (shp, kind) in shapes.items():
        if kind in ("filter", "bias", "save"):
            net.vars[k].assign(np.asarray(p[k], np.float32))
    pt, _, kinds = O.to_torch_params(a, p, B, torch.float64, requires_grad=False)
    l2_val = float(O.l2_term(pt, kinds))
    for stage in range(G.N_STAGES):
        wav, ids, mel = G.train_inputs(name, stage)
        kw = {} if mel is None else dict(mel=torch.as_tensor(mel).cuda())
        dw, di = net._prepare_inputs(torch.as_tensor(G.raw_audio(wav)) if raw else torch.as_tensor(wav), torch.as_tensor(ids))
        if raw:
            assert np.array_equal(dw.cpu().numpy(), wav)
        logits = eng.forward(dw, di, want_logits=True, **kw)
        eng.backward()
        torch.cuda.synchronize()
        s = "s%d_" % stage
        lg = logits.float().cpu().numpy()[:, ::G.logit_stride(name), :]
        ref = gold[s + "logits"]
        assert np.abs(lg - ref).max() <= 0.03 * max(1.0, np.abs(ref).max()), (stage, np.abs(lg - ref).max())
        st = eng.read_stats()
        # total loss = mean xent + l2_factor * L2 (tmodel.py:261); the engine reports the parts
        xent_ref = float(gold[s + "loss"]) - l2 * l2_val
        assert abs(st["xent_sum"] / max(st["n_valid"], 1) - xent_ref) <= 5e-3 * xent_ref, (stage, st, xent_ref)
        worst = {}
        for k, (shp, kind) in shapes.items():
            if kind not in ("filter", "bias"):
                continue
            g = eng.view(k, eng.grads)[tuple(slice(0, d) for d in shp)].double().cpu().numpy() / max(st["n_valid"], 1)
            if kind == "filter":
                g = g + l2 * np.asarray(p[k], np.float64)  # the engine's gradient is unnormalised and has no L2 term (tmodel.py:250-261)
            nrm = float(gold[s + "gradnorm_" + k])
            if nrm == 0:
                assert np.abs(g).max() == 0, k
                continue
            got, refs = G.sample_of(g, G.sample_limit(name)), gold[s + "grad_" + k]
            worst[k] = float(np.sqrt(((got - refs) ** 2).sum()) / max(np.sqrt((refs ** 2).sum()), 1e-30))
        util.record("reference_run_%s_stage%d" % (name, stage),
                    dict(logits_max_abs_err=float(np.abs(lg - ref).max()), grad_rel_err_max=max(worst.values()),
                         grad_rel_err_median=float(np.median(list(worst.values())))))
        # against fp64 at ~100 loss positions: the same bound as tests/test_gpu_train.py::test_gradients_match_oracle (the
        # bf16 forward dominates; the same-rounding oracle, which the CPU test above ties to the reference run at 1e-9,
        # is held to 6 % there)
        # (arch5: 50 layers at 512 loss positions -- the bf16 residual stream decorrelates over the depth, DESIGN.md section 4;
        # tests/test_gpu_full.py holds the same stack to 8 % at the benchmark's stage length)
        deep = a.n_layers >= 30
        bad = {k: v for k, v in worst.items() if v > (0.4 if deep else 0.2)}
        assert not bad, (stage, bad)
        assert float(np.median(list(worst.values()))) <= (0.2 if deep else 0.1)
        save_worst = 0.0
        for li, ((b, bl), dil) in enumerate(zip(a.layer_ids(), a.dilations())):
            key = "SAVE_%d_%d_%d" % (dil, b, bl)
            sv = G.save_sample(eng.save_view(li).float().cpu().numpy()[:, :, :arch["n_res"]])
            err = float(np.abs(sv - gold[s + key]).max() / max(1.0, np.abs(gold[s + key]).max()))
            save_worst = max(save_worst, err)
            # rows of the bf16 residual stream: one rounding (2^-9) per layer on top of the layers below
            assert err <= (0.05 if deep else 0.02), (stage, key, err)
        util.record("reference_run_%s_stage%d_save" % (name, stage), dict(save_rel_err_max=save_worst))
