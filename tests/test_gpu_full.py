"""GPU parity at the sizes and depths that are benchmarked (BASELINE.json configs[1], configs[0]-shaped, configs[4]).

Why these tests look the way they do (measured with the CPU oracle): the per-tensor gradient
error of ANY bf16-operand forward against the fp64 oracle is a statistical quantity -- every position contributes an
independent rounding error while the true gradient adds up coherently -- so it falls as 1/sqrt(valid positions):
3x10 stack, emulated-rounding oracle vs fp64: 600 positions 7.9 % median / 13.7 % max, 2 400: 4.4 / 10.3 %, 9 600:
2.4 / 8.9 %, 38 400: 1.0 / 5.0 %.  (Carrying the residual stream in fp32 does NOT change these figures: 7.9 / 13.8 % at
600 positions.)  The deep stacks are therefore compared with the oracle at the FULL stage length the benchmark uses,
where 8 % per tensor is a meaningful bound, and a defect inside one deep layer is caught by the per-layer isolation
tests below (every layer's forward and backward against the fp64 statement of that ONE layer, fed with the kernel's
own stash: <= 1e-2).
"""
import numpy as np
import pytest
import torch

from oracle import wavenet_oracle as O
from tests import util

pytestmark = pytest.mark.gpu

GRAD_TOL = 0.08   # per-tensor rel-L2, vs the fp64 oracle AND vs the same-rounding oracle, at full stage length
WIDE4 = dict(util.WIDE, n_blocks=4, n_block_layers=10)   # BASELINE configs[4]


def _engine(arch, B):
    from lb_wavenet_b200.engine import TrainEngine
    return TrainEngine(arch, B)


# sizes: >= 30 000 valid positions each, so that the statistical bf16 floor of the worst tensor (PRE: a row of the
# embedding table only sees the positions that carry its code) stays under the bound -- measured on B200 at 4 x 8192
# (21 881 valid): arch1 8.3 %, at 1 x 8192 wide (5 596 valid): 11.9 %, at 4 x 16384 3x10 (43 608 valid): 4.6 %
@pytest.mark.parametrize("arch,B,T", [(util.CLASSIC, 4, 16384), (util.C1, 8, 8192), (WIDE4, 6, 8192), (util.ARCH5, 8, 8192)],
                         ids=["configs1_3x10_T16384", "configs0_arch1_T8192", "configs4_wide_T8192", "arch5_gc_lc_T8192"])
def test_full_stage_length_vs_oracle(lib, arch, B, T):
    """Logits, loss statistics and EVERY gradient tensor against the fp64 oracle (autograd) and the same-rounding
    oracle (hand-written backward) at the benchmark's stage length: 128 tiles per slot through the dynamic tile
    scheduler, dil = 512 across many tiles, TMA zero fill at the stage end."""
    a = util.oracle_arch(arch)
    p = util.scaled_params(a, B, 21)
    wav, ids = util.synth_batch(B, T, max(arch["n_gc_category"], 3), 22)
    mel = util.synth_mel(B, T, a, 23, wav)   # None unless the architecture has local conditioning (reference par/arch5.json)
    eng = _engine(arch, B)
    eng.load_state(p)
    logits = eng.forward(torch.as_tensor(wav).cuda(), torch.as_tensor(ids).cuda(), want_logits=True,
                         mel=None if mel is None else torch.as_tensor(mel).cuda()).cpu().numpy()
    eng.backward()
    torch.cuda.synchronize()
    st = eng.read_stats()
    grads, L, fwd = O.train_step_autograd(a, p, wav, ids, 0.0, torch.float64, mel=mel)
    pt, save, kinds = O.to_torch_params(a, p, B, torch.float64, requires_grad=False)
    w, i = torch.as_tensor(wav).long(), torch.as_tensor(ids).long()
    gem, info = O.train_backward_manual(a, pt, save, w, i, torch.float64, emulate_bf16=True,
                                        mel=None if mel is None else torch.as_tensor(mel, dtype=torch.float64))
    lg_ex, lg_em = fwd.logits.detach().numpy(), info["fwd"].logits.numpy()
    assert st["n_valid"] == L.n_valid == info["n_valid"] and L.n_valid > 0.5 * B * T
    rec = dict(logits_rel_vs_emulated=util.rel_err(logits, lg_em), logits_rel_vs_fp64=util.rel_err(logits, lg_ex),
               logits_maxabs_vs_emulated=float(np.abs(logits - lg_em).max()))
    assert rec["logits_rel_vs_emulated"] <= 3e-2 and rec["logits_rel_vs_fp64"] <= 3e-2, rec
    assert abs(st["xent_sum"] - info["xent_sum"]) <= 2e-3 * abs(info["xent_sum"])
    assert abs(st["xent_sum"] - float(L.xent_sum)) <= 2e-3 * abs(float(L.xent_sum))
    Lg = O.loss_fn(a, torch.as_tensor(logits, dtype=torch.float64), w, i, pt, kinds, 0.0)
    assert st["diff_sum"] == Lg.diff_sum
    # SAVE: bit-exact copies of the last dil rows of [old SAVE ; x_l] of the kernel's own x_l
    new_state = eng.export_state()
    for l in (0, 5, 9, a.n_layers - 1):
        s = eng.reg.saves[l]
        xl = eng.debug_read(0, l).cpu().numpy()
        assert np.array_equal(new_state[s.name], xl[:, T - s.dil:, :]), s.name
    vs_em, vs_ex = {}, {}
    for name in eng.reg.params:
        g = eng.view(name, eng.grads).cpu().numpy()
        ref = grads[name] * L.n_valid
        if np.abs(ref).max() == 0:
            assert np.abs(g).max() == 0, name
            continue
        vs_em[name] = util.rel_err(g, gem[name].numpy())
        vs_ex[name] = util.rel_err(g, ref)
    worst_em, worst_ex = max(vs_em, key=vs_em.get), max(vs_ex, key=vs_ex.get)
    rec.update(max_vs_emulated=vs_em[worst_em], worst_vs_emulated=worst_em,
               median_vs_emulated=float(np.median(list(vs_em.values()))),
               max_vs_fp64=vs_ex[worst_ex], worst_vs_fp64=worst_ex, median_vs_fp64=float(np.median(list(vs_ex.values()))),
               n_tensors=len(vs_ex), n_valid=L.n_valid)
    util.record("full_T_parity_R%d_S%d_L%d_B%d_T%d" % (arch["n_res"], arch["n_skip"], a.n_layers, B, T), rec)
    # 30- and 40-layer stacks: every tensor within GRAD_TOL.  The 50-layer stacks (arch1, arch5) have tensors whose bf16
    # FLOOR sits above it whatever the amount of data -- the CPU oracle evaluated with the same rounding points is itself
    # that far from fp64: PRE, whose gradient crosses all 50 layers (8.9 % at 8 x 8192, 7.7 % at
    # 16 x 8192), and at random initialisation every local-conditioning tensor (15.6 % median: the mel frames explain
    # nothing yet, the LC gradient is an incoherent sum, so its relative error IS the per-element error of dv and does not
    # average out over positions).  For a tensor above GRAD_TOL the bound is therefore the floor itself: the kernel may be
    # no further from fp64 than 1.3 x what the CPU evaluation of the same numerics contract is (+ 1 %), and no further from
    # that evaluation than two independent realisations of the same rounding noise are (1.5 x).  The LC arithmetic itself
    # is checked layer by layer at 1e-2 in tests/test_gpu_lc.py.
    floor = {k: util.rel_err(gem[k].numpy(), grads[k] * L.n_valid) for k in vs_ex}
    deep50 = a.n_layers >= 50
    bad = {k: (v, floor[k]) for k, v in vs_ex.items()
           if v > GRAD_TOL and not (deep50 and v <= 1.3 * floor[k] + 0.01)}
    assert not bad, ("vs fp64 oracle (error, floor of the bf16 contract)", bad)
    bad = {k: (v, floor[k]) for k, v in vs_em.items()
           if v > GRAD_TOL and not (deep50 and v <= 1.5 * floor[k] + 0.01)}
    assert not bad, ("vs emulated oracle (error, floor of the bf16 contract)", bad)
    non_lc = [v for k, v in vs_ex.items() if not k.startswith("LC_")]
    assert float(np.median(non_lc)) <= 0.03
    rec["n_above_tol"] = sum(v > GRAD_TOL for v in vs_ex.values())
    util.record("full_T_floor_R%d_S%d_L%d_B%d_T%d" % (arch["n_res"], arch["n_skip"], a.n_layers, B, T),
                dict(n_above_tol=rec["n_above_tol"], floor_median=float(np.median(list(floor.values()))),
                     floor_max=max(floor.values()), floor_worst=max(floor, key=floor.get)))


@pytest.mark.parametrize("arch,B,T", [(util.CLASSIC, 2, 640), (util.C1, 2, 384), (util.WIDE_DEEP, 1, 1088),
                                      (util.TINY_NOBIAS, 2, 200)],
                         ids=["3x10", "arch1_5x10_gc", "wide_1x10", "tiny_nobias"])
def test_every_layer_in_isolation(lib, arch, B, T):
    """One layer at a time, at every depth: the kernel's z_l, x_{l+1} (forward) and dx_l, filter and bias
    gradients (backward, issued phase by phase) against the fp64 statement of that single layer (oracle.layer_single,
    reference tmodel.py:117-184,325) evaluated on the KERNEL's own inputs of the layer -- x_l, the skip-path dz_l,
    dx_{l+1}.  No error compounds across layers, so the bound is the bf16 output rounding plus tanh.approx: 1e-2 rel-L2
    (a dropped tap, bias or conditioning term in any single layer is a > 10 % error)."""
    a = util.oracle_arch(arch)
    p = util.scaled_params(a, B, 31)
    wav, ids = util.synth_batch(B, T, max(arch["n_gc_category"], 3), 32)
    eng = _engine(arch, B)
    eng.load_state(p)
    eng.forward(torch.as_tensor(wav).cuda(), torch.as_tensor(ids).cuda())
    torch.cuda.synchronize()
    pt = {k: torch.tensor(np.asarray(v), dtype=torch.float64) for k, v in p.items()
          if np.asarray(v).dtype.kind == "f" and not k.startswith("SAVE")}
    it = torch.as_tensor(ids).long()
    L = a.n_layers
    dils = a.dilations()
    rd = lambda what, l: eng.debug_read(what, l).double().cpu()
    x = [rd(0, l) for l in range(L)]
    z = [rd(1, l) for l in range(L)]
    full = [torch.cat([torch.tensor(p[eng.reg.saves[l].name], dtype=torch.float64), x[l]], dim=1) for l in range(L)]
    worst = {}

    def chk(key, got, ref, tol=1e-2):
        e = util.rel_err(got.numpy(), ref.numpy())
        worst[key.split(":")[0]] = max(worst.get(key.split(":")[0], 0.0), e)
        assert e <= tol, (key, e)

    for l in range(L):
        o = O.layer_single(a, pt, l, full[l], it)
        chk("z:%d" % l, z[l], o["z"])
        if l + 1 < L:
            chk("x_next:%d" % l, x[l + 1], o["x_next"])
    # backward, one phase at a time
    eng.backward_phases(0, 1)
    dz_skip = [rd(6, l) for l in range(L)]   # every plane is final after the post-net backward
    dx_next = torch.zeros(B, T, arch["n_res"], dtype=torch.float64)
    for l in reversed(range(L)):
        eng.backward_phases(L - l, L - l + 1)
        torch.cuda.synchronize()
        o = O.layer_single(a, pt, l, full[l], it, dz_skip[l], dx_next)
        dx_next = rd(7, l)   # what the next kernel consumes: dx_l, bf16
        chk("dx:%d" % l, dx_next, o["dx"])
        sfx = "%d_%d" % a.layer_ids()[l]
        for nm in ("SIGNAL", "GATE", "RESIDUAL", "SIGNAL_BIAS", "GATE_BIAS", "RESIDUAL_BIAS"):
            key = "%s_%s" % (nm, sfx)
            if key not in eng.reg.params:
                continue
            g = eng.view(key, eng.grads).double().cpu()
            if float(o[key].abs().max()) == 0:   # RESIDUAL of the last layer
                assert float(g.abs().max()) == 0, key
                continue
            chk("%s:%d" % (nm, l), g, o[key])
    util.record("layer_isolation_R%d_L%d_B%d_T%d" % (arch["n_res"], L, B, T), worst)


def test_four_stage_trajectory_vs_oracle(lib):
    """K = 4 consecutive stages of forward + backward + TF-Adam with the D-separation state carried from stage to stage
    (reference train.py:218-252 loop over tmodel.py:165's SAVE assignment), against 4 oracle steps (fp64 autograd +
    adam_tf_step, SAVE carried the same way).  lr = 1e-3, l2 = 1e-3 as in par/par1.json."""
    arch, B, T, K = util.CLASSIC_SHALLOW, 4, 2048, 4
    lr, l2 = 1e-3, 1e-3
    a = util.oracle_arch(arch)
    p = util.scaled_params(a, B, 51)
    wav, ids = util.synth_batch(B, T * K, 3, 52)
    eng = _engine(arch, B)
    eng.load_state(p)
    po = {k: np.asarray(v, np.float64).copy() for k, v in p.items()}
    shapes = O.param_shapes(a, B)
    train_keys = [k for k, (_, kind) in shapes.items() if kind in ("filter", "bias")]
    m = {k: np.zeros_like(po[k]) for k in train_keys}
    v = {k: np.zeros_like(po[k]) for k in train_keys}
    losses = []
    for k in range(K):
        sl = slice(k * T, (k + 1) * T)
        w_k, i_k = np.ascontiguousarray(wav[:, sl]), np.ascontiguousarray(ids[:, sl])
        eng.forward(torch.as_tensor(w_k).cuda(), torch.as_tensor(i_k).cuda())
        eng.l2_loss()
        eng.backward()
        st = eng.read_stats()
        eng.adam(k + 1, lr, l2)
        torch.cuda.synchronize()
        grads, Lr, fwd = O.train_step_autograd(a, po, w_k, i_k, l2, torch.float64)
        assert st["n_valid"] == Lr.n_valid
        loss_gpu = st["xent_sum"] / st["n_valid"] + l2 * st["l2"]
        # the loss falls by ~0.04 per stage here (6.149, 6.110, 6.072, 6.028 in the oracle): 1e-3 relative resolves it
        assert abs(loss_gpu - float(Lr.total)) <= 1e-3 * abs(float(Lr.total)), (k, loss_gpu, float(Lr.total))
        losses.append((loss_gpu, float(Lr.total)))
        for name in train_keys:
            po[name], m[name], v[name] = O.adam_tf_step(po[name], grads[name], m[name], v[name], k + 1, lr)
        for li, s in enumerate(eng.reg.saves):   # carry the oracle's SAVE (fp64, unrounded) forward
            po[s.name] = fwd.new_save[li].numpy()
        state = eng.export_state()
        # SAVE rows are bf16 copies of x_l rows: compare with the oracle's rows to bf16 resolution + stack error
        for li, s in enumerate(eng.reg.saves):
            assert util.rel_err(state[s.name], po[s.name]) <= 2e-2, (k, s.name)
    # after 4 Adam steps every weight moved by <= 4 * lr; what the kernel did to them agrees with the oracle
    err = {}
    for name in train_keys:
        moved_ref = po[name] - np.asarray(p[name], np.float64)
        moved_gpu = state[name].astype(np.float64) - np.asarray(p[name], np.float64)
        if np.abs(moved_ref).max() == 0:
            continue
        err[name] = util.rel_err(moved_gpu, moved_ref)
    util.record("trajectory_4_stages", dict(max=max(err.values()), median=float(np.median(list(err.values()))),
                                            losses_gpu_vs_oracle=losses))
    assert losses[0][1] - losses[-1][1] > 0.05   # the trajectory actually moves
    # Adam's first steps are sign-like (|update| ~ lr whatever |g|): an element whose gradient is near zero flips its
    # whole step on a 1 % gradient error, so the bound is on each tensor's displacement w(4) - w(0), not per element.
    # Calibration: the same-rounding CPU oracle against the fp64 oracle gives median 2.7 %, max 11.6 % (a bias vector).
    assert float(np.median(list(err.values()))) <= 0.06, err
    bad = {k: e for k, e in err.items() if e > 0.2}
    assert not bad, bad


def test_sixty_step_loss_curve_vs_oracle(lib):
    """Training quality against the fp64 reference graph over a longer run: 60 consecutive stages of the 3x10
    stack (forward + backward + TF-Adam, SAVE carried from stage to stage, lr = 1e-3, l2 = 1e-3) on the GPU and in the
    oracle, each evolving its OWN weights from the same start.  The two loss curves must stay together (1.5e-2 relative
    at every step) while the loss itself falls from 8.26 to 5.01 -- a systematic gradient or optimiser defect bends the
    GPU curve away within a few steps (10 % less progress is a 6 % deviation at the end).  Calibration: the CPU
    evaluation of the same bf16 contract drifts 2.3e-3 from fp64 over these 60 steps; the GPU run is not repeatable to
    the last bit (fp32 atomics in the weight gradients) and the two trajectories amplify that: four runs of ONE binary
    measured 5.4e-3, 6.6e-3, 7.2e-3 and 9.4e-3 (round 2's first bound, 6e-3, sat inside that spread)."""
    arch, B, T, K = util.CLASSIC, 4, 1024, 60
    lr, l2 = 1e-3, 1e-3
    a = util.oracle_arch(arch)
    p = util.scaled_params(a, B, 81)
    batches = [util.synth_batch(B, T, 3, 82 + k) for k in range(K)]   # ~70 % valid positions in every stage
    eng = _engine(arch, B)
    eng.load_state(p)
    po = {k: np.asarray(v, np.float64).copy() for k, v in p.items()}
    shapes = O.param_shapes(a, B)
    keys = [k for k, (_, kind) in shapes.items() if kind in ("filter", "bias")]
    m = {k: np.zeros_like(po[k]) for k in keys}
    v = {k: np.zeros_like(po[k]) for k in keys}
    gpu, ref = [], []
    for k in range(K):
        w_k, i_k = batches[k]
        eng.forward(torch.as_tensor(w_k).cuda(), torch.as_tensor(i_k).cuda())
        eng.l2_loss()
        eng.backward()
        st = eng.read_stats()
        eng.adam(k + 1, lr, l2)
        gpu.append(st["xent_sum"] / max(1, st["n_valid"]) + l2 * st["l2"])
        grads, Lr, fwd = O.train_step_autograd(a, po, w_k, i_k, l2, torch.float64)
        assert st["n_valid"] == Lr.n_valid
        ref.append(float(Lr.total.detach()))
        for name in keys:
            po[name], m[name], v[name] = O.adam_tf_step(po[name], grads[name], m[name], v[name], k + 1, lr)
        for li, s in enumerate(eng.reg.saves):
            po[s.name] = fwd.new_save[li].numpy()
    gpu, ref = np.array(gpu), np.array(ref)
    dev = np.abs(gpu - ref) / np.abs(ref)
    util.record("loss_curve_60_steps_3x10", dict(max_rel_dev=float(dev.max()), at=int(dev.argmax()), first=float(ref[0]),
                                                 last=float(ref[-1]), gpu_last=float(gpu[-1])))
    assert ref[0] - ref[-1] > 2.0, (ref[0], ref[-1])      # the run learns: the loss falls by >> the tolerance
    assert dev.max() <= 1.5e-2, (int(dev.argmax()), float(dev.max()))
