"""GPU parity for the primitive ops: mu-law (bit-exact), sampler (bit-exact), tcgen05 self-test."""
import numpy as np
import pytest
import torch

from oracle import wavenet_oracle as O

pytestmark = pytest.mark.gpu


def test_mu_encode_bit_exact_exhaustive(lib):
    """reference ops.py:23-28: every int16 PCM value / 32768, plus random and edge floats."""
    from lb_wavenet_b200 import _lib
    xs = [np.arange(-32768, 32768, dtype=np.float32) / np.float32(32768.0),
          np.random.default_rng(0).uniform(-1, 1, 200000).astype(np.float32),
          np.array([-1.0, 1.0, 0.0, -0.0, 1e-30, -1e-30, np.nextafter(np.float32(1), np.float32(0))], np.float32)]
    thr = O.mu_encode_thresholds()
    xs.append(np.concatenate([thr, np.nextafter(thr, np.float32(-2)), np.nextafter(thr, np.float32(2))]).clip(-1, 1))
    x = np.concatenate(xs).astype(np.float32)
    dx = torch.as_tensor(x).cuda()
    dq = torch.empty(x.size, dtype=torch.int32, device="cuda")
    _lib.check(lib.wn_mu_encode(dx.data_ptr(), dq.data_ptr(), x.size, _lib.cur_stream()))
    assert np.array_equal(dq.cpu().numpy(), O.mu_encode_np(x))
    # empty input is a no-op
    _lib.check(lib.wn_mu_encode(None, None, 0, _lib.cur_stream()))


def test_mu_decode_bit_exact(lib):
    from lb_wavenet_b200 import _lib
    q = torch.arange(256, dtype=torch.int32, device="cuda")
    out = torch.empty(256, dtype=torch.float32, device="cuda")
    _lib.check(lib.wn_mu_decode(q.data_ptr(), out.data_ptr(), 256, _lib.cur_stream()))
    assert np.array_equal(out.cpu().numpy().view(np.uint32), O.mu_decode_np(np.arange(256)).view(np.uint32))


def test_sampler_bit_exact(lib):
    """tf.multinomial replacement (reference imodel.py:179): same indices as the oracle sampler on
    identical logits, for peaked, flat and extreme rows."""
    from lb_wavenet_b200 import _lib
    rng = np.random.default_rng(1)
    n = 4096
    lg = (rng.normal(size=(n, 256)) * rng.choice([0.1, 1.0, 5.0, 30.0], size=(n, 1))).astype(np.float32)
    lg[0] = 0.0
    lg[1] = -1e4
    lg[1, 77] = 0.0
    lg[2, :] = 100.0
    d = torch.as_tensor(lg).cuda()
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    for seed, step in ((0, 0), (1234567890123, 5), (2 ** 63 + 5, 2 ** 33 + 1)):
        _lib.check(lib.wn_sample_logits(d.data_ptr(), n, seed, step, out.data_ptr(), _lib.cur_stream()))
        u = O.sampler_uniform(seed, np.full(n, step, np.uint64), np.arange(n))
        ref = O.sample_from_logits(lg, u)
        assert np.array_equal(out.cpu().numpy(), ref)
    assert out.cpu().numpy()[1] == 77


@pytest.mark.parametrize("sw", [128, 64, 32, -128, -64, -32])
@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (256, 256, 256), (200, 32, 128), (128, 96, 960)])
def test_umma_selftest_gemm(lib, sw, M, N, K):
    """TMA + tcgen05.mma + TMEM pipeline (the building blocks of the training kernels) against a
    plain fp32 matmul of the same bf16 operands.  Negative sw: thread-written swizzled tiles."""
    from lb_wavenet_b200 import _lib
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    B = torch.randn(N, K, generator=g).to(torch.bfloat16).cuda()
    Cd = torch.zeros(M, N, dtype=torch.float32, device="cuda")
    _lib.check(lib.wn_selftest_umma_gemm(A.data_ptr(), B.data_ptr(), Cd.data_ptr(), M, N, K, sw, _lib.cur_stream()))
    torch.cuda.synchronize()
    ref = A.float().cpu() @ B.float().cpu().T
    err = (Cd.cpu() - ref).abs().max().item()
    assert err <= 1e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (256, 256, 256), (512, 128, 192), (300, 64, 960)])
def test_umma_selftest_gemm_cta_pair(lib, M, N, K):
    """The same contraction on CTA pairs (tcgen05 cta_group::2): a cluster of two CTAs shares one copy of every B block
    (each loads half of its rows), the leader's M = 256 MMAs read both CTAs' shared memory and fill both CTAs' tensor
    memory.  The recipe check for moving the post-net GEMMs onto pairs."""
    from lb_wavenet_b200 import _lib
    g = torch.Generator().manual_seed(M * 5 + N * 3 + K)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16).cuda()
    B = torch.randn(N, K, generator=g).to(torch.bfloat16).cuda()
    Cd = torch.zeros(M, N, dtype=torch.float32, device="cuda")
    _lib.check(lib.wn_selftest_umma_gemm_pair(A.data_ptr(), B.data_ptr(), Cd.data_ptr(), M, N, K, _lib.cur_stream()))
    torch.cuda.synchronize()
    ref = A.float().cpu() @ B.float().cpu().T
    err = (Cd.cpu() - ref).abs().max().item()
    assert err <= 1e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("M,N,K", [(128, 256, 640), (960, 256, 1000), (64, 64, 200), (256, 96, 4096), (200, 16, 77)])
def test_umma_selftest_gemm_tn_mn_major(lib, M, N, K):
    """The split-K weight-gradient kernel (both operands MN-major, TMA SW128 panels, fp32 atomics) against
    a plain fp32 matmul: C = A^T B with A [K, M], B [K, N] row-major; ragged M, N, K included."""
    from lb_wavenet_b200 import _lib
    g = torch.Generator().manual_seed(M + 3 * N + 7 * K)
    A = torch.randn(K, M, generator=g).to(torch.bfloat16).cuda()
    B = torch.randn(K, N, generator=g).to(torch.bfloat16).cuda()
    Cd = torch.zeros(M, N, dtype=torch.float32, device="cuda")
    _lib.check(lib.wn_selftest_umma_gemm_tn(A.data_ptr(), B.data_ptr(), Cd.data_ptr(), M, N, K, _lib.cur_stream()))
    torch.cuda.synchronize()
    ref = A.float().cpu().T @ B.float().cpu()
    err = (Cd.cpu() - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err
