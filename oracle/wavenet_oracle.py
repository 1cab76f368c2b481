"""CPU ORACLE for the lb-wavenet dilated-causal-convolution hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``lb_wavenet_b200``) never does; it fails loudly
when the CUDA library is missing.

PARITY PIN: the reference (hrbigelow/lb-wavenet) ships no golden vectors, no known-answer tests and no
assertions for this path, and its arithmetic lives in TensorFlow 1.x (version unpinned by the reference, not
installable here: Python 3.12, no network).  Round 2 pins this file against OUTPUTS OF THE REFERENCE'S OWN
PYTHON RUN IN THIS CONTAINER: tests/golden/make_reference_vectors.py imports tmodel.py, arch.py, ops.py and
data.py unmodified from the reference tree and executes them on oracle/tf1_shim (an eager torch stand-in for
the ~50 `tf.*` calls they make); tests/test_reference_vectors.py holds train_forward / loss_fn / the autograd
backward / the data path here to those vectors at 1e-9 (fp64) over 4 architectures (plain, global conditioning,
local conditioning, the tiny bias-free arch2 shapes) x 2 stages with the SAVE state carried by the reference.
What that pins is the reference's graph construction (names, shapes, concat / slice / dilation / mask /
normalisation logic, op order); what it cannot pin is TensorFlow's kernels themselves -- the shim restates the
documented semantics of each op (tests/test_tf1_shim.py checks them against explicit loops).  The generator
(imodel.py) cannot be run that way: its constructor no longer matches arch.py's (imodel.py:26-37 vs arch.py:31-47),
under any TensorFlow; it stays pinned through "teacher-forced generator == training forward".
Besides that (tests/test_oracle_pins.py):
  * the README known-answer diagram (reference README.md:70-85, images/wavenet_influence.png),
  * two structurally different statements of the dilated conv (explicit two-tap form vs
    torch.nn.functional.conv1d) -- reference tmodel.py:143-144 vs imodel.py:107-108,
  * staged == whole-sequence (reference README.md:16-21),
  * teacher-forced incremental generator == training forward (reference tests.py:1,7-11 intent),
  * autograd == hand-written backward == fp64 finite differences,
  * mu-law fixed points (reference ops.py:23-39), Philox4x32-10 Random123 known answers.

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# Architecture / variable registry   (reference arch.py:6-22, 85-103, 112-142)
# --------------------------------------------------------------------------------------


@dataclass
class Arch:
    n_blocks: int
    n_block_layers: int
    n_quant: int
    n_res: int
    n_dil: int
    n_skip: int
    n_post: int
    n_gc_embed: int = 0
    n_gc_category: int = 0
    use_bias: bool = True
    # local conditioning (reference tmodel.py:15-17, arch.py:75-80): mel channels, conditioning channels, strides
    n_lc_in: int = 0
    n_lc_out: int = 0
    lc_upsample: Tuple[int, ...] = ()

    def has_lc(self) -> bool:
        # reference arch.py:108-109 use_lc_input
        return self.n_lc_out > 0

    def lc_hop(self) -> int:
        # reference train.py:129-130
        h = 1
        for s_ in self.lc_upsample:
            h *= int(s_)
        return h if self.has_lc() else 1

    @property
    def n_layers(self) -> int:
        return self.n_blocks * self.n_block_layers

    def dilations(self) -> List[int]:
        # reference tmodel.py:313-318: dil = 2**bl inside every block
        return [2 ** bl for _ in range(self.n_blocks) for bl in range(self.n_block_layers)]

    def layer_ids(self) -> List[Tuple[int, int]]:
        return [(b, bl) for b in range(self.n_blocks) for bl in range(self.n_block_layers)]

    def recep_field(self) -> int:
        # reference tmodel.py:50-51
        return self.n_blocks * sum(2 ** l for l in range(self.n_block_layers))

    def has_gc(self) -> bool:
        # reference arch.py:105-106
        return self.n_gc_embed > 0


def param_shapes(a: Arch, batch_sz: int) -> "OrderedDict[str, Tuple[Tuple[int, ...], str]]":
    """Serial names (checkpoint keys) -> (shape, kind) in graph-construction order.

    kind is 'filter' (trainable, L2-regularised), 'bias' (trainable, no L2),
    'save' (non-trainable D-separation state) or 'counter' (int32 scalar).
    Names: reference arch.py:126,142 ('_'.join([NAME(+'_BIAS'), *indices])).
    Shapes: reference arch.py:85-103.  Order: reference tmodel.py:292-328.
    """
    R, D, S, P, Q, G = a.n_res, a.n_dil, a.n_skip, a.n_post, a.n_quant, a.n_gc_embed
    out: "OrderedDict[str, Tuple[Tuple[int, ...], str]]" = OrderedDict()

    def add(name, shape, kind):
        out[name] = (tuple(shape), kind)

    if a.has_gc():
        add("GC_EMBED", (a.n_gc_category + 1, G), "filter")  # tmodel.py:92-94
    add("PRE", (Q, R), "filter")  # tmodel.py:96
    if a.use_bias:
        add("PRE_BIAS", (R,), "bias")  # tmodel.py:98-100
    if a.has_lc():  # tmodel.py:68-83 (called from build right after _preprocess, tmodel.py:307-311); shape arch.py:75-80
        for i, s_ in enumerate(a.lc_upsample):
            add("LC_UPSAMPLE_{}".format(i), (int(s_), a.n_lc_out, a.n_lc_in if i == 0 else a.n_lc_out), "filter")
    for (b, bl), dil in zip(a.layer_ids(), a.dilations()):
        sfx = "{}_{}".format(b, bl)
        add("SAVE_{}_{}".format(dil, sfx), (batch_sz, dil, R), "save")  # tmodel.py:123-124
        for nm in ("SIGNAL", "GATE"):  # tmodel.py:136-148
            add("{}_{}".format(nm, sfx), (2, R, D), "filter")
            if a.use_bias:
                add("{}_BIAS_{}".format(nm, sfx), (D,), "bias")
        if a.has_gc():  # tmodel.py:150-154
            add("GC_SIGNAL_{}".format(sfx), (G, D), "filter")
            add("GC_GATE_{}".format(sfx), (G, D), "filter")
        if a.has_lc():  # tmodel.py:156-160; shape arch.py:96-97
            add("LC_SIGNAL_{}".format(sfx), (a.n_lc_out, D), "filter")
            add("LC_GATE_{}".format(sfx), (a.n_lc_out, D), "filter")
        add("RESIDUAL_{}".format(sfx), (D, R), "filter")  # tmodel.py:171-181
        if a.use_bias:
            add("RESIDUAL_BIAS_{}".format(sfx), (R,), "bias")
        add("SKIP_{}".format(sfx), (D, S), "filter")
        if a.use_bias:
            add("SKIP_BIAS_{}".format(sfx), (S,), "bias")
    add("POST1", (S, P), "filter")  # tmodel.py:196-201
    if a.use_bias:
        add("POST1_BIAS", (P,), "bias")
    add("POST2", (P, Q), "filter")  # tmodel.py:205-210
    if a.use_bias:
        add("POST2_BIAS", (Q,), "bias")
    add("GLOBAL_STEP", (), "counter")  # tmodel.py:223-226
    add("VALID_SAMPLES", (), "counter")
    return out


def xavier_bound(shape: Sequence[int]) -> float:
    """tf.contrib.layers.xavier_initializer_conv2d (uniform) bound, reference arch.py:63.

    fan_in = shape[-2]*prod(shape[:-2]), fan_out = shape[-1]*prod(shape[:-2]).
    """
    recept = 1
    for s in shape[:-2]:
        recept *= s
    fan_in = shape[-2] * recept
    fan_out = shape[-1] * recept
    return math.sqrt(6.0 / (fan_in + fan_out))


def init_params(a: Arch, batch_sz: int, seed: int = 0, save_init: str = "xavier",
                bias_scale: float = 0.0) -> Dict[str, np.ndarray]:
    """Xavier-uniform filters, zero biases (reference arch.py:63-64,125-134).

    SAVE variables get the default (Xavier) initialiser in the reference because
    tmodel.py:123-124 passes no initializer; save_init='zero' is available for tests.
    bias_scale>0 draws non-zero biases (tests only; makes bias paths observable).
    """
    rng = np.random.default_rng(seed)
    p: Dict[str, np.ndarray] = {}
    for name, (shape, kind) in param_shapes(a, batch_sz).items():
        if kind == "filter":
            bnd = xavier_bound(shape)
            p[name] = rng.uniform(-bnd, bnd, size=shape).astype(np.float32)
        elif kind == "bias":
            p[name] = (rng.uniform(-1, 1, size=shape) * bias_scale).astype(np.float32)
        elif kind == "save":
            if save_init == "zero":
                p[name] = np.zeros(shape, np.float32)
            else:
                bnd = xavier_bound(shape)
                p[name] = rng.uniform(-bnd, bnd, size=shape).astype(np.float32)
        else:
            p[name] = np.zeros((), np.int32)
    return p


# --------------------------------------------------------------------------------------
# mu-law   (reference ops.py:23-39, the float32 numpy twins are normative)
# --------------------------------------------------------------------------------------


def mu_encode_np(x: np.ndarray, n_quanta: int = 256) -> np.ndarray:
    """reference ops.py:23-28, evaluated in float32 at every step.

    (numpy-1.x value-based casting, which the reference was written against, keeps every
    intermediate float32 for float32 input; numpy>=2 would promote the division by the
    float64 scalar log1p(mu) -- so the float32 casts are spelled out here.)
    """
    x = np.asarray(x, np.float32)
    mu = np.float32(n_quanta - 1)
    log1p_mu = np.float32(np.log1p(np.float64(n_quanta - 1)))
    amp = np.sign(x) * np.log1p(mu * np.abs(x)) / log1p_mu
    quant = (amp + np.float32(1)) * np.float32(0.5) * mu + np.float32(0.5)
    return quant.astype(np.int32)


def mu_decode_np(quant: np.ndarray, n_quanta: int = 256) -> np.ndarray:
    """reference ops.py:31-39, float32."""
    mu = np.float32(n_quanta - 1)
    qf = np.asarray(quant).astype(np.float32)
    inv_mu = np.float32(1.0 / (n_quanta - 1))
    a = (np.float32(2) * qf - np.float32(1)) * inv_mu - np.float32(1)
    x = np.sign(a) * (np.power(np.float32(1) + mu, np.fabs(a)) - np.float32(1)) * inv_mu
    return x.astype(np.float32)


def mu_encode_thresholds(n_quanta: int = 256) -> np.ndarray:
    """thr[q-1] = smallest float32 x in [-1, 1] with mu_encode_np(x) >= q, q = 1..n_quanta-1.

    mu_encode_np is monotone non-decreasing in x, so encode(x) == #{q : thr[q-1] <= x}
    for every float32 x in [-1, 1]: a table-driven device encoder is bit-exact by
    construction.  Found by bisection over the ordered float32 bit patterns.
    """

    def key_to_f32(k: np.ndarray) -> np.ndarray:
        # order-preserving map int64 key -> float32 (negative floats reversed)
        # k >= 0 -> bit pattern k; k < 0: -1 -> -0.0 (0x80000000), -2 -> 0x80000001, ...
        k = np.asarray(k, np.int64)
        neg = (np.int64(0x80000000) + (-k - 1)) & np.int64(0xFFFFFFFF)
        bits = np.where(k >= 0, k, neg).astype(np.uint32)
        return bits.view(np.float32)

    lo_key = -(int(np.float32(1.0).view(np.uint32)) + 1)  # -1.0
    hi_key = int(np.float32(1.0).view(np.uint32))  # +1.0
    thr = np.empty(n_quanta - 1, np.float32)
    for q in range(1, n_quanta):
        lo, hi = lo_key, hi_key  # invariant: enc(hi) >= q ; find smallest
        assert mu_encode_np(key_to_f32(np.array([hi])))[0] >= q
        if mu_encode_np(key_to_f32(np.array([lo])))[0] >= q:
            thr[q - 1] = key_to_f32(np.array([lo]))[0]
            continue
        while hi - lo > 1:
            mid = (lo + hi) // 2
            if mu_encode_np(key_to_f32(np.array([mid])))[0] >= q:
                hi = mid
            else:
                lo = mid
        thr[q - 1] = key_to_f32(np.array([hi]))[0]
    return thr


# --------------------------------------------------------------------------------------
# bf16 emulation helpers (numerics contract of the CUDA path, DESIGN.md "Numerics")
# --------------------------------------------------------------------------------------


def bf16_round(t: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even to bfloat16, returned in the tensor's own dtype."""
    return t.to(torch.float32).to(torch.bfloat16).to(t.dtype)


# Rounding sites of the CUDA path's contract, by name; tools/error_budget.py switches single sites off (ROUND_OFF) to
# attribute the bf16 error of a gradient to where it comes from.  Empty = the contract as it is (every test uses that).
ROUND_OFF: set = set()


def _maybe(t: torch.Tensor, emulate: bool, site: str = "") -> torch.Tensor:
    return bf16_round(t) if (emulate and site not in ROUND_OFF) else t


# --------------------------------------------------------------------------------------
# Training forward   (reference tmodel.py:292-328)
# --------------------------------------------------------------------------------------


@dataclass
class FwdResult:
    logits: torch.Tensor  # [B, T, Q]
    new_save: List[torch.Tensor]  # per layer [B, dil, R]
    xs: List[torch.Tensor] = field(default_factory=list)  # layer inputs x_l  [B,T,R]
    zs: List[torch.Tensor] = field(default_factory=list)  # gated outputs z_l [B,T,D]
    skip_sum: Optional[torch.Tensor] = None
    x_out: Optional[torch.Tensor] = None  # residual stream after the last layer (unused by the loss)


def _t(p, name, dtype):
    v = p[name]
    if isinstance(v, torch.Tensor):
        return v.to(dtype)
    return torch.as_tensor(np.asarray(v), dtype=dtype)


def lc_upsample(a: Arch, p: Dict[str, torch.Tensor], mel: torch.Tensor, emulate_bf16: bool = False,
                impl: str = "gemm", keep: Optional[list] = None) -> torch.Tensor:
    """reference tmodel.py:68-83 _preprocess_lc: a chain of tf.contrib.nn.conv1d_transpose(lc, filt, out_shape, stride)
    with filt [width, out_channels, in_channels] (arch.py:75-80) and width == stride == lc_upsample[i], no bias, no
    non-linearity: out[b, t*s + k, o] = sum_c in[b, t, c] * filt[k, o, c]  ('SAME' and 'VALID' agree when width == stride).
    impl 'gemm' = that formula as one matmul per level; 'conv_transpose' = torch.nn.functional.conv_transpose1d, a
    structurally different statement of the same op.  emulate_bf16: the CUDA path's rounding points (mel, every level's
    output and the filters are bf16 operands).  keep: list that receives every level's input (for the backward)."""
    lc = _maybe(mel, emulate_bf16, "lc")
    for i, s_ in enumerate(a.lc_upsample):
        filt = _maybe(p["LC_UPSAMPLE_{}".format(i)], emulate_bf16, "w")  # [s, n_out, n_in]
        if keep is not None:
            keep.append(lc)
        B, Ti, _ = lc.shape
        if impl == "gemm":
            out = torch.einsum("btc,koc->btko", lc, filt).reshape(B, Ti * int(s_), filt.shape[1])
        else:
            w = filt.permute(2, 1, 0)  # conv_transpose1d weight: [in_channels, out_channels, width]
            out = torch.nn.functional.conv_transpose1d(lc.transpose(1, 2), w, stride=int(s_)).transpose(1, 2)
        lc = _maybe(out, emulate_bf16, "lc")
    return lc


def train_forward(a: Arch, p: Dict[str, torch.Tensor], save: List[torch.Tensor],
                  wav: torch.Tensor, ids: torch.Tensor, dtype=torch.float64,
                  emulate_bf16: bool = False, conv_impl: str = "taps",
                  keep: bool = False, mel: Optional[torch.Tensor] = None) -> FwdResult:
    """Forward of WaveNetTrain.build (reference tmodel.py:292-328).

    p: name -> tensor already in ``dtype`` (requires_grad as the caller wishes).
    save: per-layer SAVE tensors [B, dil, R] (reference tmodel.py:123-124).
    wav: int64 [B, T] mu-law codes; ids: int64 [B, T] (0 == invalid).
    emulate_bf16: apply the CUDA path's rounding points (operands of every contraction and
    the stored activations are bf16; accumulation stays in ``dtype``).
    conv_impl: 'taps' = explicit full[t]*W[0] + full[t+dil]*W[1] (imodel.py:107-108 form),
               'conv1d' = torch conv1d with dilation on [SAVE;cur] (tmodel.py:143-144 form).
    """
    B, T = wav.shape
    em = emulate_bf16

    def W(name):  # contraction operand (bf16 in the CUDA path)
        return _maybe(p[name], em, "w")

    # tmodel.py:53-66 one-hot (out-of-range index -> zero row) ; tmodel.py:96-100 PRE 1x1.
    # one_hot @ PRE == row gather, which is how the CUDA path does it (fp32 table).
    valid = ((wav >= 0) & (wav < a.n_quant)).to(dtype).unsqueeze(-1)
    cur = p["PRE"][wav.clamp(0, a.n_quant - 1)] * valid
    if a.use_bias:
        cur = cur + p["PRE_BIAS"]
    cur = _maybe(cur, em, "x")

    if a.has_gc():
        gathered = p["GC_EMBED"][ids]  # tmodel.py:112  [B,T,G]
    lc_up = None
    if a.has_lc():  # tmodel.py:307-311; mel [B, T / hop, n_lc_in]
        lc_up = lc_upsample(a, p, mel.to(dtype), em)
        assert lc_up.shape[1] == T, (lc_up.shape, T)

    new_save, xs, zs = [], [], []
    skp_sum = None
    for li, ((b, bl), dil) in enumerate(zip(a.layer_ids(), a.dilations())):
        sfx = "{}_{}".format(b, bl)
        full = torch.cat([save[li].to(dtype), cur], dim=1)  # tmodel.py:127  [B, dil+T, R]
        v = {}
        for nm in ("SIGNAL", "GATE"):
            filt = W("{}_{}".format(nm, sfx))  # [2, R, D]
            if conv_impl == "taps":
                vv = full[:, :T, :] @ filt[0] + full[:, dil:dil + T, :] @ filt[1]
            else:
                w = filt.permute(2, 1, 0).contiguous()  # [D, R, 2]
                vv = torch.nn.functional.conv1d(full.transpose(1, 2), w, dilation=dil).transpose(1, 2)
            if a.use_bias:
                vv = vv + p["{}_BIAS_{}".format(nm, sfx)]  # tmodel.py:145-148
            if a.has_gc():  # tmodel.py:150-154 (fp32 table in the CUDA path, no bf16 rounding)
                vv = vv + gathered @ p["GC_{}_{}".format(nm, sfx)]
            if a.has_lc():  # tmodel.py:156-160 (the CUDA path stores the projection as a bf16 plane)
                vv = vv + _maybe(lc_up @ W("LC_{}_{}".format(nm, sfx)), em, "lc")
            v[nm] = vv
        new_save.append(full[:, full.shape[1] - dil:, :].detach())  # tmodel.py:165
        z = torch.tanh(v["SIGNAL"]) * torch.sigmoid(v["GATE"])  # tmodel.py:167
        z = _maybe(z, em, "z")
        sig = z @ W("RESIDUAL_" + sfx)  # tmodel.py:171-181
        skp = z @ W("SKIP_" + sfx)
        if a.use_bias:
            sig = sig + p["RESIDUAL_BIAS_" + sfx]
            skp = skp + p["SKIP_BIAS_" + sfx]
        skp_sum = skp if skp_sum is None else skp_sum + skp  # tmodel.py:321-324
        if keep:
            xs.append(cur)
            zs.append(z)
        cur = _maybe(cur + sig, em, "x")  # tmodel.py:325

    # tmodel.py:187-215 (softmax output unused)
    h1 = _maybe(torch.relu(skp_sum), em, "h")
    d1 = h1 @ W("POST1")
    if a.use_bias:
        d1 = d1 + p["POST1_BIAS"]
    h2 = _maybe(torch.relu(d1), em, "h")
    logits = h2 @ W("POST2")
    if a.use_bias:
        logits = logits + p["POST2_BIAS"]
    return FwdResult(logits=logits, new_save=new_save, xs=xs, zs=zs, skip_sum=skp_sum, x_out=cur)


@dataclass
class LossResult:
    total: torch.Tensor
    xent_mean: torch.Tensor
    xent_sum: torch.Tensor
    n_valid: int
    avg_diff: int
    diff_sum: int
    l2: torch.Tensor
    mask: torch.Tensor


def l2_term(p: Dict[str, torch.Tensor], kinds: Dict[str, str]) -> torch.Tensor:
    """reference tmodel.py:250-258: sum of tf.nn.l2_loss (= 0.5*sum(v**2)) over trainable
    variables whose key does not contain 'BIAS'."""
    tot = None
    for k, v in p.items():
        if kinds[k] == "filter" and "BIAS" not in k:
            t = 0.5 * (v * v).sum()
            tot = t if tot is None else tot + t
    return tot


def loss_fn(a: Arch, logits: torch.Tensor, wav: torch.Tensor, ids: torch.Tensor,
            p: Dict[str, torch.Tensor], kinds: Dict[str, str], l2_factor: float) -> LossResult:
    """reference tmodel.py:228-261."""
    labels = wav[:, 1:]  # tmodel.py:230 (logits[t] predicts input[t+1])
    lg = logits[:, :-1, :]  # tmodel.py:231
    mask_i = (ids[:, 1:] != 0)  # tmodel.py:232
    mask = mask_i.to(logits.dtype)
    lse = torch.logsumexp(lg, dim=2)
    picked = lg.gather(2, labels.clamp(0, a.n_quant - 1).unsqueeze(-1)).squeeze(-1)
    # tmodel.py:64: tf.one_hot of an out-of-range code is an ALL-ZERO row, so softmax_cross_entropy_with_logits_v2
    # (tmodel.py:235) returns -sum(0 * log_softmax) = 0 there, argmax(label) (tmodel.py:240) is 0, and the op's
    # registered gradient is grad_loss * (softmax - labels) = softmax (TF's fused kernel hands back softmax - labels
    # whatever the labels sum to).  The zero-valued second term carries exactly that gradient through autograd.
    lab_ok = (labels >= 0) & (labels < a.n_quant)
    sm = torch.softmax(lg, dim=2).detach()
    xent_bad = (sm * (lg - lg.detach())).sum(dim=2)
    xent = torch.where(lab_ok, lse - picked, xent_bad) * mask  # tmodel.py:235-237
    lab_arg = torch.where(lab_ok, labels, torch.zeros_like(labels))
    diffs = (lab_arg - lg.argmax(dim=2)).abs() * mask_i  # tmodel.py:240-241 (int32 in TF)
    diff_sum = int(diffs.sum().item())
    avg_diff = diff_sum // diffs.numel() if diffs.numel() else 0  # tmodel.py:242 integer reduce_mean
    n_valid = int(mask_i.sum().item())  # tmodel.py:244
    xent_sum = xent.sum()
    xent_mean = xent_sum / n_valid if n_valid != 0 else xent_sum * 0.0  # tmodel.py:246-249
    l2 = l2_term(p, kinds)
    total = xent_mean + l2_factor * l2  # tmodel.py:261
    return LossResult(total, xent_mean, xent_sum, n_valid, avg_diff, diff_sum, l2, mask)


def to_torch_params(a: Arch, p_np: Dict[str, np.ndarray], batch_sz: int, dtype=torch.float64,
                    requires_grad: bool = True):
    """Split a numpy param dict into (trainable torch dict, save list, kinds)."""
    shapes = param_shapes(a, batch_sz)
    kinds = {k: kind for k, (_, kind) in shapes.items()}
    p, save = OrderedDict(), []
    for k, (shape, kind) in shapes.items():
        if kind in ("filter", "bias"):
            t = torch.tensor(np.asarray(p_np[k]), dtype=dtype)
            t.requires_grad_(requires_grad)
            p[k] = t
        elif kind == "save":
            save.append(torch.tensor(np.asarray(p_np[k]), dtype=dtype))
    return p, save, kinds


def train_step_autograd(a: Arch, p_np: Dict[str, np.ndarray], wav: np.ndarray, ids: np.ndarray,
                        l2_factor: float, dtype=torch.float64, emulate_bf16: bool = False, mel=None):
    """One forward+backward through torch autograd (the oracle for tmodel.grad_var,
    reference tmodel.py:354-358).  Returns (grads dict name->np, LossResult, FwdResult)."""
    B = wav.shape[0]
    p, save, kinds = to_torch_params(a, p_np, B, dtype)
    w = torch.as_tensor(np.asarray(wav), dtype=torch.int64)
    i = torch.as_tensor(np.asarray(ids), dtype=torch.int64)
    mt = None if mel is None else torch.as_tensor(np.asarray(mel), dtype=dtype)
    fwd = train_forward(a, p, save, w, i, dtype, emulate_bf16, keep=True, mel=mt)
    L = loss_fn(a, fwd.logits, w, i, p, kinds, l2_factor)
    L.total.backward()
    grads = {k: (v.grad.detach().numpy().copy() if v.grad is not None else np.zeros(tuple(v.shape)))
             for k, v in p.items()}
    return grads, L, fwd


# --------------------------------------------------------------------------------------
# Hand-written backward (second statement of the gradient; also carries the CUDA path's
# backward rounding points when emulate_bf16=True).  Unnormalised: the returned gradients
# are d(sum of masked xent)/d(param); divide by n_valid and add l2_factor*w to obtain the
# gradient of the reference's total loss (tmodel.py:246-261).
# --------------------------------------------------------------------------------------


def train_backward_manual(a: Arch, p: Dict[str, torch.Tensor], save: List[torch.Tensor],
                          wav: torch.Tensor, ids: torch.Tensor, dtype=torch.float64,
                          emulate_bf16: bool = False, mel: Optional[torch.Tensor] = None):
    em = emulate_bf16
    B, T = wav.shape
    with torch.no_grad():
        pd = {k: v.detach().to(dtype) for k, v in p.items()}

        def W(name):
            return _maybe(pd[name], em, "w")

        fwd = train_forward(a, pd, save, wav, ids, dtype, em, keep=True, mel=mel)
        lc_in: list = []
        lc_up, dlc_up = None, None
        if a.has_lc():
            lc_up = lc_upsample(a, pd, mel.to(dtype), em, keep=lc_in)
            dlc_up = torch.zeros_like(lc_up)
        # recompute post-net intermediates
        h1 = _maybe(torch.relu(fwd.skip_sum), em, "h")
        d1 = h1 @ W("POST1") + (pd["POST1_BIAS"] if a.use_bias else 0)
        h2 = _maybe(torch.relu(d1), em, "h")
        logits = fwd.logits
        # loss gradient wrt logits (unnormalised), tmodel.py:230-237
        mask = torch.zeros(B, T, dtype=dtype)
        mask[:, :-1] = (ids[:, 1:] != 0).to(dtype)
        labels = torch.zeros(B, T, dtype=torch.int64)
        labels[:, :-1] = wav[:, 1:]
        sm = torch.softmax(logits, dim=2)
        lab_ok = ((labels >= 0) & (labels < a.n_quant))
        onehot = torch.nn.functional.one_hot(labels.clamp(0, a.n_quant - 1), a.n_quant).to(dtype)
        onehot = onehot * lab_ok.unsqueeze(-1).to(dtype)  # out-of-range code: all-zero one-hot row (tmodel.py:64)
        dlog = _maybe((sm - onehot) * mask.unsqueeze(-1), em, "dlog")
        g: Dict[str, torch.Tensor] = {}
        flat = lambda t: t.reshape(-1, t.shape[-1])
        g["POST2"] = flat(h2).T @ flat(dlog)
        if a.use_bias:
            g["POST2_BIAS"] = flat(dlog).sum(0)
        dp1 = _maybe((dlog @ W("POST2").T) * (h2 > 0).to(dtype), em, "dpost")
        g["POST1"] = flat(h1).T @ flat(dp1)
        if a.use_bias:
            g["POST1_BIAS"] = flat(dp1).sum(0)
        dskip = _maybe((dp1 @ W("POST1").T) * (h1 > 0).to(dtype), em, "dpost")  # [B,T,S]
        dx = torch.zeros(B, T, a.n_res, dtype=dtype)  # grad wrt x_{l+1}; zero after last layer
        if a.has_gc():
            gathered = pd["GC_EMBED"][ids]
            g["GC_EMBED"] = torch.zeros_like(pd["GC_EMBED"])
        for li in reversed(range(a.n_layers)):
            (b, bl), dil = a.layer_ids()[li], a.dilations()[li]
            sfx = "{}_{}".format(b, bl)
            x, z = fwd.xs[li], fwd.zs[li]
            full = torch.cat([save[li].to(dtype), x], dim=1)
            xa, xb = full[:, :T, :], full[:, dil:dil + T, :]  # x[t-dil], x[t]
            v = {}
            for nm in ("SIGNAL", "GATE"):
                filt = W("{}_{}".format(nm, sfx))
                vv = xa @ filt[0] + xb @ filt[1]
                if a.use_bias:
                    vv = vv + pd["{}_BIAS_{}".format(nm, sfx)]
                if a.has_gc():
                    vv = vv + gathered @ pd["GC_{}_{}".format(nm, sfx)]
                if a.has_lc():
                    vv = vv + _maybe(lc_up @ W("LC_{}_{}".format(nm, sfx)), em, "lc")
                v[nm] = vv
            th, sg = torch.tanh(v["SIGNAL"]), torch.sigmoid(v["GATE"])
            g["SKIP_" + sfx] = flat(z).T @ flat(dskip)
            g["RESIDUAL_" + sfx] = flat(z).T @ flat(dx)
            if a.use_bias:
                g["SKIP_BIAS_" + sfx] = flat(dskip).sum(0)
                g["RESIDUAL_BIAS_" + sfx] = flat(dx).sum(0)
            dz_skip = _maybe(dskip @ W("SKIP_" + sfx).T, em, "dz")
            dz = dz_skip + dx @ W("RESIDUAL_" + sfx).T
            dvs = _maybe(dz * sg * (1 - th * th), em, "dv")
            dvg = _maybe(dz * th * sg * (1 - sg), em, "dv")
            for nm, dv in (("SIGNAL", dvs), ("GATE", dvg)):
                gw = torch.stack([flat(xa).T @ flat(dv), flat(xb).T @ flat(dv)])
                g["{}_{}".format(nm, sfx)] = gw
                if a.use_bias:
                    g["{}_BIAS_{}".format(nm, sfx)] = flat(dv).sum(0)
                if a.has_gc():
                    g["GC_{}_{}".format(nm, sfx)] = flat(gathered).T @ flat(dv)
                    dgath = dv @ pd["GC_{}_{}".format(nm, sfx)].T  # [B,T,G]
                    g["GC_EMBED"].index_add_(0, ids.reshape(-1), flat(dgath))
                if a.has_lc():  # dv (already a bf16 tile in the CUDA path) is the gradient wrt the conditioning plane
                    g["LC_{}_{}".format(nm, sfx)] = flat(lc_up).T @ flat(dv)
                    dlc_up += dv @ W("LC_{}_{}".format(nm, sfx)).T
            # data gradient: x_l[t] feeds v[t] through W[1] and v[t+dil] through W[0];
            # rows that fall in the SAVE prefix receive none (truncated at the stage boundary)
            Ws, Wg = W("SIGNAL_" + sfx), W("GATE_" + sfx)
            cur_part = dvs @ Ws[1].T + dvg @ Wg[1].T
            old_part = dvs @ Ws[0].T + dvg @ Wg[0].T  # gradient for full[t] , t in [0,T)
            # CUDA path rounding points: the data gradient is stored in split form
            # dx_l[t] = Y_l[t] + P0_l[t+dil] with Y_l = bf16(dx_{l+1} + cur_part), P0_l = bf16(old_part)
            # and the consuming layer merges the two into one bf16 tile: dx_l = bf16(Y_l[t] + P0_l[t+dil])
            dxl = _maybe(dx + cur_part, em, "dx")
            if T > dil:
                dxl[:, :T - dil, :] += _maybe(old_part, em, "dx")[:, dil:, :]
            dx = _maybe(dxl, em, "dx") if li > 0 else dxl  # layer 0's gradient feeds the PRE gather in split form
        if a.has_lc():  # the upsampling chain in reverse (tmodel.py:68-83)
            d = _maybe(dlc_up, em, "lc")
            for i in reversed(range(len(a.lc_upsample))):
                s_ = int(a.lc_upsample[i])
                filt = W("LC_UPSAMPLE_{}".format(i))  # [s, n_out, n_in]
                xin = lc_in[i]  # [B, Ti, n_in]
                dv_ = d.reshape(xin.shape[0], xin.shape[1], s_, filt.shape[1])  # [B, Ti, k, o]
                g["LC_UPSAMPLE_{}".format(i)] = torch.einsum("btko,btc->koc", dv_, xin)
                if i > 0:
                    d = _maybe(torch.einsum("btko,koc->btc", dv_, filt), em, "lc")
        # PRE gather backward
        g["PRE"] = torch.zeros_like(pd["PRE"])
        okay = ((wav >= 0) & (wav < a.n_quant)).reshape(-1)
        g["PRE"].index_add_(0, wav.reshape(-1).clamp(0, a.n_quant - 1)[okay], flat(dx)[okay])
        if a.use_bias:
            g["PRE_BIAS"] = flat(dx).sum(0)
        n_valid = int((ids[:, 1:] != 0).sum().item())
        lse = torch.logsumexp(logits[:, :-1], dim=2)
        picked = logits[:, :-1].gather(2, wav[:, 1:].clamp(0, a.n_quant - 1).unsqueeze(-1)).squeeze(-1)
        xent_sum = ((lse - picked) * lab_ok[:, :-1].to(dtype) * mask[:, :-1]).sum()
    return g, dict(n_valid=n_valid, xent_sum=float(xent_sum), fwd=fwd)


# --------------------------------------------------------------------------------------
# ONE layer in isolation (tests feed it the CUDA path's own stash for layer l, so a defect in a single deep layer is
# not hidden behind the rounding noise of the 30-50 layers around it).  Same statements as train_forward /
# train_backward_manual above: reference tmodel.py:117-168 (_dilated_conv), :171-184 (_chan_reduce), :325 (residual add).
# --------------------------------------------------------------------------------------


def layer_single(a: Arch, p: Dict[str, torch.Tensor], li: int, x_full: torch.Tensor, ids: Optional[torch.Tensor] = None,
                 dz_skip: Optional[torch.Tensor] = None, dx_next: Optional[torch.Tensor] = None,
                 round_weights: bool = True, lc_up: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """x_full [B, dil+T, R] = [SAVE ; x_l] (tmodel.py:127).  Forward: z_l, x_{l+1}.  With dz_skip [B,T,D] (gradient wrt
    z_l through the skip branch) and dx_next [B,T,R] (gradient wrt x_{l+1}; zeros after the last layer) also the
    unnormalised backward of the layer: filter / bias gradients, and the data gradient in the split form the CUDA path
    stores, dx_l[t] = Y[t] + P0[t + dil]  (Y = dx_next + dv . W[1]^T, P0 = dv . W[0]^T; rows t + dil >= T get no P0:
    the gradient stops at the stage boundary, SAVE is a variable, tmodel.py:123-124,165)."""
    (b, bl), dil = a.layer_ids()[li], a.dilations()[li]
    sfx = "{}_{}".format(b, bl)
    dt = x_full.dtype
    T = x_full.shape[1] - dil

    def W(name):
        w = p[name].to(dt)
        return bf16_round(w) if round_weights else w

    xa, xb = x_full[:, :T, :], x_full[:, dil:dil + T, :]  # x[t-dil], x[t]  (tmodel.py:143-144)
    v = {}
    gathered = p["GC_EMBED"].to(dt)[ids] if a.has_gc() else None
    for nm in ("SIGNAL", "GATE"):
        filt = W("{}_{}".format(nm, sfx))
        vv = xa @ filt[0] + xb @ filt[1]
        if a.use_bias:
            vv = vv + p["{}_BIAS_{}".format(nm, sfx)].to(dt)
        if a.has_gc():
            vv = vv + gathered @ p["GC_{}_{}".format(nm, sfx)].to(dt)
        if lc_up is not None:  # tmodel.py:156-160; lc_up [B, T, n_lc_out]
            vv = vv + lc_up.to(dt) @ W("LC_{}_{}".format(nm, sfx))
        v[nm] = vv
    th, sg = torch.tanh(v["SIGNAL"]), torch.sigmoid(v["GATE"])
    z = th * sg
    sig = z @ W("RESIDUAL_" + sfx)
    if a.use_bias:
        sig = sig + p["RESIDUAL_BIAS_" + sfx].to(dt)
    out = dict(z=z, x_next=xb + sig)
    if dz_skip is None:
        return out
    flat = lambda t: t.reshape(-1, t.shape[-1])
    dz = dz_skip + dx_next @ W("RESIDUAL_" + sfx).T
    dvs = dz * sg * (1 - th * th)
    dvg = dz * th * sg * (1 - sg)
    out["dv"] = torch.cat([dvs, dvg], dim=2)  # gradient wrt the pre-activations == wrt the conditioning terms added to them
    if lc_up is not None:
        out["LC_SIGNAL_" + sfx] = flat(lc_up.to(dt)).T @ flat(dvs)
        out["LC_GATE_" + sfx] = flat(lc_up.to(dt)).T @ flat(dvg)
    out["RESIDUAL_" + sfx] = flat(z).T @ flat(dx_next)
    if a.use_bias:
        out["RESIDUAL_BIAS_" + sfx] = flat(dx_next).sum(0)
    for nm, dv in (("SIGNAL", dvs), ("GATE", dvg)):
        out["{}_{}".format(nm, sfx)] = torch.stack([flat(xa).T @ flat(dv), flat(xb).T @ flat(dv)])
        if a.use_bias:
            out["{}_BIAS_{}".format(nm, sfx)] = flat(dv).sum(0)
    Ws, Wg = W("SIGNAL_" + sfx), W("GATE_" + sfx)
    out["Y"] = dx_next + dvs @ Ws[1].T + dvg @ Wg[1].T
    out["P0"] = dvs @ Ws[0].T + dvg @ Wg[0].T
    dx = out["Y"].clone()
    if T > dil:
        dx[:, :T - dil, :] += out["P0"][:, dil:, :]
    out["dx"] = dx
    return out


# --------------------------------------------------------------------------------------
# TF Adam   (reference train.py:178,186 -> tf.train.AdamOptimizer defaults)
# --------------------------------------------------------------------------------------


def adam_tf_step(w: np.ndarray, g: np.ndarray, m: np.ndarray, v: np.ndarray, t: int, lr: float,
                 beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8):
    """TF1 AdamOptimizer update (epsilon-hat form):
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
    w -= lr_t * m / (sqrt(v) + eps).  t starts at 1."""
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    w = w - lr_t * m / (np.sqrt(v) + eps)
    return w, m, v


# --------------------------------------------------------------------------------------
# Counter-based sampler (replaces unseeded tf.multinomial, reference imodel.py:179;
# SURVEY quirk ledger: X).  Philox4x32-10 (Salmon et al., Random123) + a fully specified
# float32 inverse-CDF so that device and oracle agree bit for bit on identical logits.
# --------------------------------------------------------------------------------------

_PH_M0, _PH_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PH_W0, _PH_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr [...,4] uint32, key [...,2] uint32 -> [...,4] uint32."""
    c = [np.asarray(ctr[..., i], np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], np.uint64)
    k1 = np.asarray(key[..., 1], np.uint64)
    M32 = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = _PH_M0 * c[0]
        p1 = _PH_M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & M32
        hi1, lo1 = p1 >> np.uint64(32), p1 & M32
        c = [(hi1 ^ c[1] ^ k0) & M32, lo1, (hi0 ^ c[3] ^ k1) & M32, lo0]
        k0 = (k0 + np.uint64(_PH_W0)) & M32
        k1 = (k1 + np.uint64(_PH_W1)) & M32
    return np.stack(c, axis=-1).astype(np.uint32)


def sampler_uniform(seed: int, step: np.ndarray, stream: np.ndarray) -> np.ndarray:
    """u in [0,1) float32 with 24 random bits: key=(seed lo, seed hi),
    counter=(step lo, step hi, stream, 0), word 0 of the Philox output."""
    step = np.asarray(step, np.uint64)
    stream = np.asarray(stream, np.uint64)
    step, stream = np.broadcast_arrays(step, stream)
    ctr = np.stack([step & np.uint64(0xFFFFFFFF), step >> np.uint64(32), stream & np.uint64(0xFFFFFFFF),
                    np.zeros_like(step)], axis=-1).astype(np.uint32)
    key = np.empty(step.shape + (2,), np.uint32)
    key[..., 0] = np.uint32(seed & 0xFFFFFFFF)
    key[..., 1] = np.uint32((seed >> 32) & 0xFFFFFFFF)
    w0 = philox4x32_10(ctr, key)[..., 0]
    return ((w0 >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)


_EXP2_COEF = [np.float32(c) for c in (
    1.0, 0.6931471805599453, 0.2402265069591007, 0.05550410866482158,
    0.009618129107628477, 0.0013333558146428443, 0.00015403530393381608)]


def det_exp(x: np.ndarray) -> np.ndarray:
    """Deterministic float32 exp(x) for x <= 0: separate RN multiplies and adds only (no fma),
    so the CUDA sampler (using __fmul_rn/__fadd_rn) reproduces it bit for bit.
    t = x*log2(e); n = rint(t); f = t-n; 2^f by a degree-6 Horner polynomial; scale by 2^n.
    Returns 0 for t < -120."""
    x = np.asarray(x, np.float32)
    t = (x * np.float32(1.4426950408889634)).astype(np.float32)
    small = t < np.float32(-120.0)
    t = np.where(small, np.float32(0), t).astype(np.float32)
    n = np.rint(t).astype(np.float32)
    f = (t - n).astype(np.float32)
    pz = np.full_like(f, _EXP2_COEF[6])
    for c in reversed(_EXP2_COEF[:6]):
        pz = (pz * f).astype(np.float32)
        pz = (pz + c).astype(np.float32)
    scale = ((n.astype(np.int32) + 127) << 23).astype(np.int32).view(np.float32)
    out = (pz * scale).astype(np.float32)
    return np.where(small, np.float32(0), out).astype(np.float32)


def sample_from_logits(logits: np.ndarray, u: np.ndarray) -> np.ndarray:
    """Inverse-CDF categorical sample per row of float32 logits [N, 256].

    Canonical evaluation order (mirrored by the CUDA sampler, one warp per row):
      e_i = det_exp(l_i - max_j l_j)
      lane k (0..31) owns entries 8k..8k+7 and forms running sums s_{k,0..7} sequentially;
      lane totals are combined by a Kogge-Stone inclusive scan (offsets 1,2,4,8,16);
      c_i = exclusive_prefix(lane) + s_{k,j};  total = inclusive prefix of lane 31;
      thr = u * total;  sample = min(#{i : c_i <= thr}, 255).
    """
    lg = np.asarray(logits, np.float32)
    N, Q = lg.shape
    assert Q == 256
    e = det_exp(lg - lg.max(axis=1, keepdims=True))
    e = e.reshape(N, 32, 8)
    s = np.empty_like(e)
    run = np.zeros((N, 32), np.float32)
    for j in range(8):
        run = (run + e[:, :, j]).astype(np.float32)
        s[:, :, j] = run
    incl = s[:, :, 7].copy()
    for d in (1, 2, 4, 8, 16):
        sh = np.zeros_like(incl)
        sh[:, d:] = incl[:, :-d]
        upd = (incl + sh).astype(np.float32)
        incl = np.where(np.arange(32)[None, :] >= d, upd, incl).astype(np.float32)
    excl = np.zeros_like(incl)
    excl[:, 1:] = incl[:, :-1]
    c = (excl[:, :, None] + s).astype(np.float32).reshape(N, 256)
    total = incl[:, 31]
    thr = (np.asarray(u, np.float32) * total).astype(np.float32)
    idx = (c <= thr[:, None]).sum(axis=1)
    return np.minimum(idx, 255).astype(np.int32)


# --------------------------------------------------------------------------------------
# Incremental generator   (reference imodel.py:61-272)
# --------------------------------------------------------------------------------------


class GenOracle:
    """State machine of WaveNetGen._loop_body (reference imodel.py:214-272).

    Deviations (SURVEY quirk ledger, all 'X'): PRE bias is added (imodel.py:75-77 omits it,
    which makes the generator inconsistent with the trainer, tmodel.py:98-100); true ring
    buffers instead of chunk-shifted lookback buffers (imodel.py:88-97,199-201 -- same values
    are read: position wpos holds x[t-dil], wpos+dil receives x[t]); every sample is emitted
    (imodel.py:205,256-258 drops the trailing partial chunk); seeded counter-based sampling
    instead of tf.multinomial (imodel.py:179).
    Preserved: the first input is the all-zero vector, i.e. PRE contributes only its bias
    (imodel.py:66-70); lookback buffers start at zero (imodel.py:88-95); teacher forcing
    replaces the fed-back sample while i < len(teacher) (imodel.py:260-267).
    """

    def __init__(self, a: Arch, p_np: Dict[str, np.ndarray], n_streams: int, dtype=torch.float64,
                 emulate_bf16: bool = False, gc_ids: Optional[np.ndarray] = None):
        self.a, self.B, self.dtype, self.em = a, n_streams, dtype, emulate_bf16
        self.p = {k: torch.tensor(np.asarray(v), dtype=dtype) for k, v in p_np.items()
                  if np.asarray(v).dtype.kind == "f" and not k.startswith("SAVE")}
        self.rings = [torch.zeros(n_streams, d, a.n_res, dtype=dtype) for d in a.dilations()]
        self._w: Dict[str, torch.Tensor] = {}
        self.t = 0
        self.cur_code = np.full(n_streams, -1, np.int64)  # -1 == all-zero input vector
        self.gc = None
        if a.has_gc():
            g = np.zeros(n_streams, np.int64) if gc_ids is None else np.asarray(gc_ids, np.int64)
            self.gc = self.p["GC_EMBED"][torch.as_tensor(g)]  # imodel.py:53-56

    def W(self, name):
        w = self._w.get(name)
        if w is None:  # the weights never change during a run: round them once
            w = self._w[name] = _maybe(self.p[name], self.em)
        return w

    def step_logits(self) -> torch.Tensor:
        """Consume self.cur_code, advance all rings one timestep, return logits [B, Q]."""
        a, p, em = self.a, self.p, self.em
        code = torch.as_tensor(self.cur_code)
        valid = (code >= 0).to(self.dtype).unsqueeze(-1)
        z = p["PRE"][code.clamp(0)] * valid  # imodel.py:73-74
        if a.use_bias:
            z = z + p["PRE_BIAS"]
        z = _maybe(z, em)
        skp_all = None
        for li, ((b, bl), dil) in enumerate(zip(a.layer_ids(), a.dilations())):
            sfx = "{}_{}".format(b, bl)
            slot = self.t % dil
            old = self.rings[li][:, slot, :].clone()  # x[t-dil]   imodel.py:107
            self.rings[li][:, slot, :] = z  # imodel.py:97
            v = {}
            for nm in ("SIGNAL", "GATE"):
                filt = self.W("{}_{}".format(nm, sfx))
                vv = old @ filt[0] + z @ filt[1]  # imodel.py:107-108
                if a.use_bias:
                    vv = vv + p["{}_BIAS_{}".format(nm, sfx)]
                if self.gc is not None:
                    vv = vv + self.gc @ p["GC_{}_{}".format(nm, sfx)]  # imodel.py:113-118
                v[nm] = vv
            dconv = _maybe(torch.tanh(v["SIGNAL"]) * torch.sigmoid(v["GATE"]), em)  # imodel.py:121
            sig = dconv @ self.W("RESIDUAL_" + sfx)
            skp = dconv @ self.W("SKIP_" + sfx)
            if a.use_bias:
                sig = sig + p["RESIDUAL_BIAS_" + sfx]
                skp = skp + p["SKIP_BIAS_" + sfx]
            skp_all = skp if skp_all is None else skp_all + skp  # imodel.py:247
            z = _maybe(z + sig, em)  # imodel.py:245
        h1 = _maybe(torch.relu(skp_all), em)  # imodel.py:145-164
        d1 = h1 @ self.W("POST1") + (p["POST1_BIAS"] if a.use_bias else 0)
        h2 = _maybe(torch.relu(d1), em)
        logits = h2 @ self.W("POST2") + (p["POST2_BIAS"] if a.use_bias else 0)
        self.t += 1
        return logits

    def run(self, n_steps: int, seed: int, teacher: Optional[np.ndarray] = None,
            return_logits: bool = False):
        """Returns codes int32 [B, n_steps] (the sampled index at every step, before teacher
        substitution, exactly what imodel.py:179-182 writes to wav_buf), optionally logits."""
        out = np.zeros((self.B, n_steps), np.int32)
        lgs = []
        n_teacher = 0 if teacher is None else len(teacher)
        streams = np.arange(self.B)
        for i in range(n_steps):
            logits = self.step_logits().to(torch.float32).numpy()
            u = sampler_uniform(seed, np.full(self.B, i), streams)
            samp = sample_from_logits(logits, u)
            out[:, i] = samp
            if return_logits:
                lgs.append(logits)
            if i < n_teacher:  # imodel.py:260-267
                self.cur_code = np.full(self.B, int(teacher[i]), np.int64)
            else:
                self.cur_code = samp.astype(np.int64)
        if return_logits:
            return out, np.stack(lgs, axis=1)
        return out


# --------------------------------------------------------------------------------------
# Data path   (reference data.py:32-37, 110-227)
# --------------------------------------------------------------------------------------


def align_slice_sz(slice_sz: int, mel_hop_sz: int) -> int:
    """reference data.py:32-37: round slice_sz UP to a multiple of mel_hop_sz."""
    if slice_sz % mel_hop_sz != 0:
        slice_sz += mel_hop_sz - (slice_sz % mel_hop_sz)
    return slice_sz


def gen_concat_slices(file_iter: Iterator[Tuple[int, int, np.ndarray]], slice_sz: int,
                      recep_field_sz: int, mel_hop_sz: int = 1):
    """One slot generator (reference data.py:110-191), python-loop restatement.

    file_iter yields (datum_count, voice_id, wav int array) and is SHARED between slots
    (data.py:211).  Yields (datum_count, wav[slice_sz] int32, ids[slice_sz] int32).
    """
    need = slice_sz
    sw: List[np.ndarray] = []
    si: List[np.ndarray] = []
    recep_bound = recep_field_sz - 1  # data.py:133
    for datum_count, vid, wav in file_iter:
        snip = len(wav) % mel_hop_sz  # data.py:141-142
        wav = wav[:-snip or None]
        wav_sz = wav.shape[0]
        if wav_sz < recep_field_sz:  # data.py:150-154
            continue
        ids = np.concatenate([np.full(recep_bound, 0, np.int32),
                              np.full(wav_sz - recep_bound, vid, np.int32)])  # data.py:156-159
        cur = 0
        while need <= wav_sz - cur:  # data.py:163-182
            sw.append(wav[cur:cur + need])
            si.append(ids[cur:cur + need])
            cur += need
            yield datum_count, np.concatenate(sw).astype(np.int32), np.concatenate(si).astype(np.int32)
            sw, si, need = [], [], slice_sz
        if cur != wav_sz:  # data.py:184-189
            sw.append(wav[cur:])
            si.append(ids[cur:])
            need -= wav_sz - cur


def gen_slice_batches(file_iter, batch_sz: int, slice_sz: int, recep_field_sz: int,
                      mel_hop_sz: int = 1):
    """reference data.py:194-227: batch_sz slot generators over one shared file iterator,
    pulled in slot order; yields (latest_file_read_count, wav[B,T], ids[B,T])."""
    gens = [gen_concat_slices(file_iter, slice_sz, recep_field_sz, mel_hop_sz) for _ in range(batch_sz)]
    while True:
        try:
            batch = [next(g) for g in gens]
        except StopIteration:
            return
        yield batch[-1][0], np.stack([b[1] for b in batch]), np.stack([b[2] for b in batch])


def shuffled_repeat_order(n_files: int, seed: int, skip: int = 0) -> Iterator[int]:
    """File-index stream of ds.repeat().shuffle(buffer_size=n_files, seed).skip(k)
    (reference data.py:246-250): a streaming shuffle buffer over the endlessly repeated
    catalog.  TF's own RNG stream cannot be reproduced without TF (quirk ledger: X); the
    buffer algorithm is the same, the generator is numpy PCG64 seeded with ``seed``.
    """
    rng = np.random.Generator(np.random.PCG64(seed & ((1 << 63) - 1)))
    buf = list(range(n_files))
    nxt = 0  # next index of the repeated stream (mod n_files)
    produced = 0
    while True:
        j = int(rng.integers(0, len(buf)))
        val = buf[j]
        buf[j] = nxt
        nxt = (nxt + 1) % n_files
        if produced >= skip:
            yield val
        produced += 1
