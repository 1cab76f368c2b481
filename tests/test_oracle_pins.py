"""What pins the CPU oracle (the reference ships no golden vectors).

Runs on CPU.  Each test names the reference statement it is derived from.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import wavenet_oracle as O
from tests import util

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _linear_stack_arch():
    # dilations [1,2,4,8]*3, one channel everywhere (reference README.md:70-85)
    return O.Arch(3, 4, 256, 1, 1, 1, 1, 0, 0, True)


def test_readme_known_answer_diagram():
    """README.md:70-85 / images/wavenet_influence.png: all filters [0.5, 0.5], input ...0 | 4096...
    Driven through the real oracle forward by making every layer linear: signal taps (+0.5, -0.5) with
    RESIDUAL = 1 gives x' = x[t] + 0.5 x[t-d] - 0.5 x[t]; the gate is held at 1 by a large bias and
    the input is scaled by eps so that tanh is linear to ~1e-9."""
    with open(os.path.join(GOLD, "readme_influence.json")) as f:
        kat = json.load(f)
    a = _linear_stack_arch()
    eps = 1e-7
    T_left, T_right = 20, 60
    p = {k: np.zeros(s, np.float64) for k, (s, kind) in O.param_shapes(a, 1).items() if kind in ("filter", "bias")}
    p["PRE"][1, 0] = 4096 * eps  # code 0 -> 0, code 1 -> 4096
    for (b, bl) in a.layer_ids():
        sfx = "%d_%d" % (b, bl)
        p["SIGNAL_" + sfx][0, 0, 0] = 0.5
        p["SIGNAL_" + sfx][1, 0, 0] = -0.5
        p["GATE_BIAS_" + sfx][0] = 40.0
        p["RESIDUAL_" + sfx][0, 0] = 1.0
    pt = {k: torch.tensor(v) for k, v in p.items()}
    save = [torch.zeros(1, d, 1, dtype=torch.float64) for d in a.dilations()]
    wav = torch.tensor([[0] * T_left + [1] * T_right])
    ids = torch.ones_like(wav)
    fwd = O.train_forward(a, pt, save, wav, ids, torch.float64, keep=True)
    rows = [x[0, :, 0].numpy() / eps for x in fwd.xs] + [fwd.x_out[0, :, 0].numpy() / eps]
    top = rows[-1]
    j = T_left  # junction
    assert np.allclose(top[j:j + len(kat["top_row_from_junction"])], kat["top_row_from_junction"], atol=1e-3)
    assert np.allclose(top[j + kat["steps_to_saturate"] - 1:j + kat["steps_to_saturate"] + 1], kat["top_row_tail"], atol=1e-3)
    assert np.allclose(top[:j], 0.0, atol=1e-6)
    assert np.allclose(rows[1][j - 1:j + 2], kat["layer1_around_junction"], atol=1e-6)
    assert np.allclose(rows[2][j - 1:j + 4], kat["layer2_around_junction"], atol=1e-6)
    # 4096 is reached exactly sum(dil) = 45 steps after the junction (true receptive field F+1 = 46)
    assert sum(a.dilations()) == kat["steps_to_saturate"] == a.recep_field()
    first_full = int(np.argmax(np.isclose(top, 4096.0, atol=1e-3)))
    assert first_full == j + 45
    # the purple outline: saved D-separation values are exactly the last dil nodes of each layer's input
    for l, d in enumerate(a.dilations()):
        assert torch.equal(fwd.new_save[l], fwd.xs[l][:, -d:, :])


def test_stage_boundary_through_junction_continues():
    """README.md:62-66: a stage boundary right after the junction still yields the values of the
    uninterrupted computation once the saved D-separation nodes are prepended."""
    a = util.oracle_arch(util.TINY)
    B, T = 2, 128
    p = util.scaled_params(a, B, 0)
    pt, save, _ = O.to_torch_params(a, p, B, torch.float64, False)
    wav, ids = util.synth_batch(B, T, 3, 1)
    w, i = torch.as_tensor(wav).long(), torch.as_tensor(ids).long()
    whole = O.train_forward(a, pt, save, w, i)
    for cuts in ([64], [1, 2, 50], [127], [5, 13, 30, 31, 100]):
        s, outs, t0 = save, [], 0
        for c in cuts + [T]:
            r = O.train_forward(a, pt, s, w[:, t0:c], i[:, t0:c])
            s, t0 = r.new_save, c
            outs.append(r.logits)
        assert torch.allclose(torch.cat(outs, 1), whole.logits, atol=1e-12), cuts
        for x, y in zip(s, whole.new_save):
            assert torch.equal(x, y)  # T < dil stages included (cuts of 1..2 samples)


def test_two_statements_of_the_dilated_conv_agree():
    """tmodel.py:143-144 (tf.nn.convolution VALID, dilation) vs imodel.py:107-108 (explicit taps):
    filt[0] multiplies x[t-dil], filt[1] multiplies x[t]."""
    a = util.oracle_arch(util.TINY_GC)
    B, T = 2, 50
    p = util.scaled_params(a, B, 2)
    pt, save, _ = O.to_torch_params(a, p, B, torch.float64, False)
    wav, ids = util.synth_batch(B, T, 11, 3)
    w, i = torch.as_tensor(wav).long(), torch.as_tensor(ids).long()
    a1 = O.train_forward(a, pt, save, w, i, conv_impl="taps")
    a2 = O.train_forward(a, pt, save, w, i, conv_impl="conv1d")
    assert torch.allclose(a1.logits, a2.logits, atol=1e-12)


@pytest.mark.parametrize("arch", [util.TINY, util.TINY_GC, util.TINY_NOBIAS])
def test_teacher_forced_generator_equals_training_forward(arch):
    """The reference's intended 'Test 1' (tests.py:1,7-11): tmodel and imodel are equivalent functions.
    Generator fed the training inputs one step late (its first input is the all-zero vector) with rings
    initialised from zero == trainer with zero SAVE and the same shifted input."""
    a = util.oracle_arch(arch)
    B, T = 3, 70
    p = util.scaled_params(a, B, 4)
    for k in p:
        if k.startswith("SAVE"):
            p[k] = np.zeros_like(p[k])
    pt, save, _ = O.to_torch_params(a, p, B, torch.float64, False)
    rng = np.random.default_rng(0)
    seq = rng.integers(0, 256, T)
    gc = np.array([1, 2, 3]) if a.has_gc() else None
    gen = O.GenOracle(a, p, B, torch.float64, gc_ids=gc)
    _, lg_gen = gen.run(T, seed=0, teacher=seq, return_logits=True)
    # trainer sees [all-zero vector, seq[0], seq[1], ...]; the all-zero vector is an out-of-range code
    wav = torch.tensor(np.tile(np.concatenate([[-1], seq[:-1]]), (B, 1)))
    ids = torch.as_tensor(np.tile(gc[:, None], (1, T))) if gc is not None else torch.ones(B, T, dtype=torch.long)
    fwd = O.train_forward(a, pt, save, wav, ids)
    assert np.allclose(lg_gen, fwd.logits.numpy(), atol=1e-10)


@pytest.mark.parametrize("arch", [util.TINY_GC, util.TINY_ASYM])
def test_backward_statements_agree_and_match_finite_differences(arch):
    """tmodel.py:354-358: autograd == hand-written backward (exact arithmetic) == central differences."""
    a = util.oracle_arch(arch)
    B, T = 2, 40
    p = util.scaled_params(a, B, 6)
    wav, ids = util.synth_batch(B, T, max(arch["n_gc_category"], 3), 7)
    grads, L, _ = O.train_step_autograd(a, p, wav, ids, 1e-3, torch.float64)
    pt, save, kinds = O.to_torch_params(a, p, B, torch.float64, False)
    w, i = torch.as_tensor(wav).long(), torch.as_tensor(ids).long()
    gm, info = O.train_backward_manual(a, pt, save, w, i, torch.float64, emulate_bf16=False)
    assert info["n_valid"] == L.n_valid > 0
    for k in grads:
        ref = grads[k]
        man = gm[k].numpy() / L.n_valid + (1e-3 * p[k].astype(np.float64) if kinds[k] == "filter" else 0)
        assert np.allclose(man, ref, rtol=1e-9, atol=1e-12), k
    rng = np.random.default_rng(0)

    def total(pp):
        t, s, kk = O.to_torch_params(a, pp, B, torch.float64, False)
        f = O.train_forward(a, t, s, w, i)
        return float(O.loss_fn(a, f.logits, w, i, t, kk, 1e-3).total)

    for k in ["PRE", "SIGNAL_0_1", "GATE_BIAS_0_2", "RESIDUAL_0_0", "SKIP_0_3", "POST1", "POST2_BIAS"] + \
            (["GC_EMBED", "GC_GATE_0_1"] if a.has_gc() else []):
        idx = tuple(rng.integers(0, s) for s in p[k].shape)
        if k == "PRE":
            idx = (int(wav[0, 3]), idx[1])
        if k == "GC_EMBED":
            idx = (int(ids[ids > 0][0]), idx[1])
        errs = []
        for h in (1e-5, 1e-6, 1e-7):  # a ReLU kink inside +-h spoils one step size, not all three
            pp = {n: v.astype(np.float64).copy() if v.dtype.kind == "f" else v for n, v in p.items()}
            pp[k][idx] += h
            up = total(pp)
            pp[k][idx] -= 2 * h
            dn = total(pp)
            fd = (up - dn) / (2 * h)
            errs.append(abs(fd - grads[k][idx]) / (1e-4 + abs(fd)))
        assert min(errs) <= 1e-4, (k, idx, errs)


def test_last_layer_residual_gets_only_the_l2_gradient():
    """tmodel.py:313-325,252-258: RESIDUAL of the last layer is unused by the loss but regularised."""
    a = util.oracle_arch(util.TINY)
    p = util.scaled_params(a, 2, 8)
    wav, ids = util.synth_batch(2, 48, 3, 9)
    g, L, _ = O.train_step_autograd(a, p, wav, ids, 0.25, torch.float64)
    last = "RESIDUAL_%d_%d" % (a.n_blocks - 1, a.n_block_layers - 1)
    assert np.allclose(g[last], 0.25 * p[last].astype(np.float64))
    assert np.allclose(g[last.replace("RESIDUAL", "RESIDUAL_BIAS")], 0.0)


def test_mask_rule_and_loss_bookkeeping():
    """data.py:133,156-159 + tmodel.py:230-249: a file of length N contributes N-F+1 non-zero ids; the loss
    counts id_mask[:,1:] != 0; logits at the last stage position are never trained; n_valid == 0 -> 0."""
    F = 10
    files = iter([(1, 5, np.arange(25) % 256), (2, 7, np.arange(9) % 256), (3, 9, np.arange(40) % 256)])
    slices = list(O.gen_concat_slices(files, 16, F))
    ids = np.concatenate([s[2] for s in slices])
    assert len(slices) == (25 + 40) // 16  # the 9-sample file is shorter than F and skipped (data.py:150-154)
    assert (ids[:25] != 0).sum() == 25 - F + 1 and (ids[:F - 1] == 0).all() and (ids[F - 1:25] == 5).all()
    assert (ids[25:25 + F - 1] == 0).all() and (ids[25 + F - 1:] == 9).all()
    a = util.oracle_arch(util.TINY)
    p = util.scaled_params(a, 1, 0)
    pt, save, kinds = O.to_torch_params(a, p, 1, torch.float64, False)
    wav = torch.as_tensor(slices[0][1][None]).long()
    idt = torch.as_tensor(slices[0][2][None]).long()
    f = O.train_forward(a, pt, save, wav, idt)
    L = O.loss_fn(a, f.logits, wav, idt, pt, kinds, 0.0)
    assert L.n_valid == int((slices[0][2][1:] != 0).sum()) == 16 - F + 1 - 0
    L0 = O.loss_fn(a, f.logits, wav, torch.zeros_like(idt), pt, kinds, 0.0)
    assert L0.n_valid == 0 and float(L0.xent_mean) == 0.0 and float(L0.total) == 0.0
    # integer reduce_mean of |argmax diff| over ALL B*(T-1) positions (tmodel.py:240-242)
    assert L.avg_diff == L.diff_sum // 15


def test_mu_law_fixed_points_monotone_thresholds():
    """ops.py:23-39: 0 -> 128, +1 -> 255, -1 -> 0, decode(128) == 0; encoder monotone; the 255-entry
    threshold table reproduces it on every probed float32 (basis of the bit-exact device encoder)."""
    assert O.mu_encode_np(np.float32([0, 1, -1, -0.0])).tolist() == [128, 255, 0, 128]
    assert O.mu_decode_np(np.array([128]))[0] == 0.0
    x = np.concatenate([np.arange(-32768, 32768, dtype=np.float32) / np.float32(32768),
                        np.random.default_rng(0).uniform(-1, 1, 100000).astype(np.float32)])
    xs = np.sort(x)
    q = O.mu_encode_np(xs)
    assert (np.diff(q) >= 0).all() and q.min() == 0 and q.max() == 255
    thr = O.mu_encode_thresholds()
    assert (np.diff(thr) > 0).all()
    assert np.array_equal((thr[None, :] <= xs[:, None]).sum(1), q)
    below = np.nextafter(thr, np.float32(-2))
    assert np.array_equal(O.mu_encode_np(thr), np.arange(1, 256))
    assert np.array_equal(O.mu_encode_np(np.clip(below, -1, 1))[1:], np.arange(1, 255))
    # round trip error is within one quantisation step everywhere (tests.py:15-24 prints this SSE)
    rt = O.mu_decode_np(O.mu_encode_np(xs))
    assert np.abs(rt - xs).max() < 0.05


def test_committed_tables_match_oracle():
    """lb_wavenet_b200/csrc/tables.inc (compiled into the library) == the oracle's tables."""
    import re
    path = os.path.join(os.path.dirname(__file__), "..", "lb_wavenet_b200", "csrc", "tables.inc")
    txt = open(path).read()

    def arr(name):
        m = re.search(name + r"[^=]*= \{([^}]*)\}", txt) or re.search(name + r" \{([^}]*)\}", txt)
        return np.array([int(v.strip().rstrip("u"), 16) for v in m.group(1).split(",")], np.uint32)

    assert np.array_equal(arr("kMuEncodeThrBits"), O.mu_encode_thresholds().view(np.uint32))
    assert np.array_equal(arr("kMuDecodeBits"), O.mu_decode_np(np.arange(256)).view(np.uint32))
    assert np.array_equal(arr("WN_EXP2_COEF_BITS"), np.array(O._EXP2_COEF, np.float32).view(np.uint32))


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for c, k, out in kat:
        got = O.philox4x32_10(np.array(c, np.uint32), np.array(k, np.uint32))
        assert got.tolist() == list(out)


def test_sampler_follows_the_softmax_distribution():
    """imodel.py:179 replacement: inverse-CDF sampling draws from softmax(logits)."""
    rng = np.random.default_rng(3)
    lg = (rng.normal(size=256) * 2).astype(np.float32)
    n = 200000
    u = O.sampler_uniform(11, np.arange(n), np.zeros(n, np.int64))
    assert 0 <= u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 5e-3
    s = O.sample_from_logits(np.tile(lg, (n, 1)), u)
    p = np.exp(lg - lg.max())
    p /= p.sum()
    freq = np.bincount(s, minlength=256) / n
    assert np.abs(freq - p).max() < 4 * np.sqrt(p.max() / n) + 1e-3
    xs = -np.abs(rng.normal(size=50000) * 20).astype(np.float32)
    ex = np.exp(xs.astype(np.float64))
    ok = xs * 1.4426950408889634 > -120
    assert np.max(np.abs(O.det_exp(xs)[ok] - ex[ok]) / ex[ok]) < 5e-6


def test_tf_adam_formula():
    """train.py:178 -> tf.train.AdamOptimizer: epsilon outside the bias correction ('epsilon hat')."""
    w, g = np.array([1.0, -2.0]), np.array([0.5, -0.25])
    w1, m1, v1 = O.adam_tf_step(w, g, np.zeros(2), np.zeros(2), 1, 0.1)
    lr_t = 0.1 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert np.allclose(w1, w - lr_t * (0.1 * g) / (np.sqrt(0.001 * g * g) + 1e-8))
    # differs from torch.optim.Adam, whose epsilon is added after bias-correcting sqrt(v)
    t = torch.tensor(w, requires_grad=True)
    opt = torch.optim.Adam([t], lr=0.1, eps=1e-8)
    t.grad = torch.tensor(g)
    opt.step()
    assert np.allclose(t.detach().numpy(), w1, atol=1e-6) and not np.array_equal(t.detach().numpy(), w1)


def test_golden_oracle_vector():
    """tests/golden/oracle_tiny_forward.npz (generated by tests/golden/make_golden.py): regression pin."""
    z = np.load(os.path.join(GOLD, "oracle_tiny_forward.npz"))
    a = O.Arch(2, 3, 256, 16, 16, 32, 32, n_gc_embed=5, n_gc_category=7)
    p = O.init_params(a, 2, seed=123, bias_scale=0.3)
    grads, L, fwd = O.train_step_autograd(a, p, z["wav"], z["ids"], 1e-3, torch.float64)
    assert np.allclose(fwd.logits.detach().numpy(), z["logits"], atol=1e-12)
    assert L.n_valid == int(z["n_valid"]) and L.diff_sum == int(z["diff_sum"])
    assert abs(float(L.total) - float(z["total"])) < 1e-12
    for k in ("PRE", "SIGNAL_1_2", "POST2"):
        assert np.allclose(grads[k], z["grad_" + k], atol=1e-12)
    assert np.array_equal(fwd.new_save[-1].numpy(), z["save_last"])


def test_bf16_gradient_error_floor():
    """Documents the precision floor quoted in DESIGN.md: with the CUDA path's rounding points the
    oracle's own gradients move by a few percent (median) at random init; backward-only rounding is
    ~10x smaller.  Guards against silently tightening/loosening the GPU tolerances."""
    a = util.oracle_arch(util.TINY)
    B, T = 3, 96
    p = util.scaled_params(a, B, 11)
    wav, ids = util.synth_batch(B, T, 3, 12)
    grads, L, _ = O.train_step_autograd(a, p, wav, ids, 0.0, torch.float64)
    pt, save, _ = O.to_torch_params(a, p, B, torch.float64, False)
    gm, _ = O.train_backward_manual(a, pt, save, torch.as_tensor(wav).long(), torch.as_tensor(ids).long(),
                                    torch.float64, emulate_bf16=True)
    errs = [util.rel_err(gm[k].numpy() / L.n_valid, grads[k]) for k in grads if np.abs(grads[k]).max() > 0]
    assert 0.01 < np.median(errs) < 0.08 and max(errs) < 0.2


def test_out_of_range_code_is_an_all_zero_one_hot_row():
    """reference tmodel.py:64 (tf.one_hot) + :230-236: an out-of-range mu-law code is an all-zero one-hot row both as an
    INPUT (PRE contributes only its bias) and as a LABEL: softmax_cross_entropy_with_logits_v2 then returns
    -sum(0 * log_softmax) = 0, tf.argmax of the zero row is 0, and the op's registered gradient is
    grad_loss * (softmax - labels) = softmax.  Both backward statements of the oracle carry exactly that."""
    arch = util.TINY
    a = util.oracle_arch(arch)
    B, T = 2, 40
    p = util.scaled_params(a, B, 3)
    wav, _ = util.synth_batch(B, T, 3, 4)
    wav = wav.copy()
    wav[0, 10], wav[1, 20] = 256, -7
    ids = np.ones((B, T), np.int32)
    pt, save, kinds = O.to_torch_params(a, p, B, torch.float64, requires_grad=False)
    w, i = torch.as_tensor(wav).long(), torch.as_tensor(ids).long()
    fwd = O.train_forward(a, pt, save, w, i, torch.float64, keep=True)
    bias = pt["PRE_BIAS"]
    assert torch.equal(fwd.xs[0][0, 10], bias) and torch.equal(fwd.xs[0][1, 20], bias)
    L = O.loss_fn(a, fwd.logits, w, i, pt, kinds, 0.0)
    # the same batch with in-range labels at those two positions differs by exactly their two cross entropies
    wav2 = wav.copy()
    wav2[0, 10], wav2[1, 20] = 5, 9
    lg = fwd.logits
    lse = torch.logsumexp(lg, dim=2)
    extra = (lse[0, 9] - lg[0, 9, 5]) + (lse[1, 19] - lg[1, 19, 9])
    L2 = O.loss_fn(a, lg, torch.as_tensor(wav2).long(), i, pt, kinds, 0.0)
    assert L.n_valid == L2.n_valid == B * (T - 1)
    assert abs(float(L2.xent_sum - L.xent_sum - extra)) < 1e-9
    am = lg.argmax(dim=2)
    assert L.diff_sum - L2.diff_sum == int(am[0, 9] + am[1, 19]) - int(abs(5 - am[0, 9]) + abs(9 - am[1, 19]))
    # gradient wrt the logits of such a row is softmax / n_valid (autograd statement) ...
    lgr = lg.clone().requires_grad_(True)
    O.loss_fn(a, lgr, w, i, pt, kinds, 0.0).total.backward()
    sm = torch.softmax(lg, dim=2)
    assert torch.allclose(lgr.grad[0, 9] * L.n_valid, sm[0, 9], atol=1e-12)
    assert torch.allclose(lgr.grad[1, 19] * L.n_valid, sm[1, 19], atol=1e-12)
    # ... and the two full backward statements agree on every parameter
    grads, La, _ = O.train_step_autograd(a, p, wav, ids, 0.0, torch.float64)
    gman, info = O.train_backward_manual(a, pt, save, w, i, torch.float64)
    assert abs(info["xent_sum"] - float(La.xent_sum)) < 1e-9
    for k in grads:
        assert np.allclose(grads[k] * La.n_valid, gman[k].numpy(), rtol=1e-9, atol=1e-10), k


def test_single_layer_statement_composes_to_the_stack():
    """oracle.layer_single (one layer of tmodel.py:117-184,325 in isolation, used by the per-layer GPU tests) chained
    over the layers reproduces train_forward, and its backward reproduces the hand-written whole-stack backward."""
    arch = util.TINY_GC
    a = util.oracle_arch(arch)
    B, T = 2, 50
    p = util.scaled_params(a, B, 13)
    wav, ids = util.synth_batch(B, T, arch["n_gc_category"], 14)
    pt, save, kinds = O.to_torch_params(a, p, B, torch.float64, requires_grad=False)
    w, i = torch.as_tensor(wav).long(), torch.as_tensor(ids).long()
    fwd = O.train_forward(a, pt, save, w, i, torch.float64, keep=True)
    x = fwd.xs[0]
    for l in range(a.n_layers):
        o = O.layer_single(a, pt, l, torch.cat([save[l], x], dim=1), i, round_weights=False)
        assert torch.allclose(o["z"], fwd.zs[l], atol=1e-12)
        x = o["x_next"]
    assert torch.allclose(x, fwd.x_out, atol=1e-12)
    # backward: rebuild every layer's gradient from layer_single, walking down from the post-net
    gman, _ = O.train_backward_manual(a, pt, save, w, i, torch.float64)
    lg = fwd.logits
    mask = torch.zeros(B, T, dtype=torch.float64)
    mask[:, :-1] = (i[:, 1:] != 0).double()
    labels = torch.zeros(B, T, dtype=torch.int64)
    labels[:, :-1] = w[:, 1:]
    dlog = (torch.softmax(lg, 2) - torch.nn.functional.one_hot(labels, a.n_quant).double()) * mask.unsqueeze(-1)
    h1 = torch.relu(fwd.skip_sum)
    h2 = torch.relu(h1 @ pt["POST1"] + pt["POST1_BIAS"])
    dskip = ((dlog @ pt["POST2"].T) * (h2 > 0) @ pt["POST1"].T) * (h1 > 0)
    dx = torch.zeros(B, T, a.n_res, dtype=torch.float64)
    for l in reversed(range(a.n_layers)):
        sfx = "%d_%d" % a.layer_ids()[l]
        o = O.layer_single(a, pt, l, torch.cat([save[l], fwd.xs[l]], dim=1), i, dskip @ pt["SKIP_" + sfx].T, dx,
                           round_weights=False)
        for nm in ("SIGNAL", "GATE", "RESIDUAL", "SIGNAL_BIAS", "GATE_BIAS", "RESIDUAL_BIAS"):
            assert torch.allclose(o["%s_%s" % (nm, sfx)], gman["%s_%s" % (nm, sfx)], rtol=1e-9, atol=1e-11), (nm, l)
        dx = o["dx"]
    assert torch.allclose(dx.reshape(-1, a.n_res).sum(0), gman["PRE_BIAS"], rtol=1e-9, atol=1e-11)


def test_local_conditioning_statements_agree():
    """reference tmodel.py:68-83 (_preprocess_lc) and :156-160: (1) the transposed convolution with width == stride as
    one matmul per level == torch's conv_transpose1d; (2) every output frame t*s + k of a level depends on input frame
    t only (hand case); (3) autograd == the hand-written backward on every LC tensor (LC_UPSAMPLE_i, LC_SIGNAL, LC_GATE)
    == fp64 finite differences of the loss."""
    arch = util.TINY_GC_LC
    a = util.oracle_arch(dict(arch, n_res=8, n_dil=8, n_skip=16, n_post=16))
    B, T = 2, 32
    p = util.scaled_params(a, B, 7)
    wav, ids = util.synth_batch(B, T, arch["n_gc_category"], 8)
    mel = util.synth_mel(B, T, a, 9)
    pt, save, kinds = O.to_torch_params(a, p, B, torch.float64, requires_grad=False)
    m = torch.as_tensor(mel, dtype=torch.float64)
    u1, u2 = O.lc_upsample(a, pt, m, impl="gemm"), O.lc_upsample(a, pt, m, impl="conv_transpose")
    assert u1.shape == (B, T, a.n_lc_out) and torch.allclose(u1, u2, atol=1e-12)
    m2 = m.clone()
    m2[0, 3] += 1.0   # frame 3 -> output rows [3 * hop, 4 * hop) of slot 0 only
    d = (O.lc_upsample(a, pt, m2) - u1).abs().sum(dim=2)
    hop = a.lc_hop()
    assert float(d[0, 3 * hop:4 * hop].min()) > 0 and float(d.sum() - d[0, 3 * hop:4 * hop].sum()) == 0.0
    grads, L, _ = O.train_step_autograd(a, p, wav, ids, 0.0, torch.float64, mel=mel)
    gm, _ = O.train_backward_manual(a, pt, save, torch.as_tensor(wav).long(), torch.as_tensor(ids).long(),
                                    torch.float64, mel=m)
    for k in grads:
        assert np.allclose(grads[k] * L.n_valid, gm[k].numpy(), rtol=1e-9, atol=1e-11), k
    # finite differences on one element of each LC tensor
    def loss_of(pp):
        q, sv, kd = O.to_torch_params(a, pp, B, torch.float64, requires_grad=False)
        f = O.train_forward(a, q, sv, torch.as_tensor(wav).long(), torch.as_tensor(ids).long(), torch.float64, mel=m)
        return float(O.loss_fn(a, f.logits, torch.as_tensor(wav).long(), torch.as_tensor(ids).long(), q, kd, 0.0).total)
    for name, idx in (("LC_UPSAMPLE_0", (1, 2, 3)), ("LC_SIGNAL_0_1", (2, 1)), ("LC_GATE_1_0", (0, 3))):
        eps = 1e-5
        pp = {k: np.array(v, np.float64) for k, v in p.items()}
        pp[name][idx] += eps
        up = loss_of(pp)
        pp[name][idx] -= 2 * eps
        dn = loss_of(pp)
        fd = (up - dn) / (2 * eps)
        assert abs(fd - grads[name][idx]) <= 1e-6 * max(1.0, abs(fd)) + 1e-9, (name, fd, grads[name][idx])
