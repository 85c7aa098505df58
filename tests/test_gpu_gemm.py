"""tcgen05/TMA GEMM against numpy (f16 inputs, f32 accumulate): exact products, so tolerance only
covers summation order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ref(a, b, mn):
    a32, b32 = a.astype(np.float32), b.astype(np.float32)
    return a32 @ (b32 if mn else b32.T)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 256), (300, 200, 192), (1500, 1280, 1280),
                                   (1500, 1500, 64), (257, 130, 240), (3000, 384, 240)])
def test_gemm_k_major(M, N, K):
    from speaksense_b200.asr import debug_gemm
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    a = rng.standard_normal((M, K)).astype(np.float16)
    b = rng.standard_normal((N, K)).astype(np.float16)
    d = debug_gemm(a, b)
    ref = _ref(a, b, False)
    err = np.abs(d - ref).max()
    assert err < 2e-3 * np.sqrt(K), (err, d[:2, :4], ref[:2, :4])


@pytest.mark.parametrize("M,N,K", [(128, 64, 64), (128, 64, 128), (1500, 64, 1536), (200, 64, 1000)])
def test_gemm_b_mn_major(M, N, K):
    from speaksense_b200.asr import debug_gemm
    rng = np.random.default_rng(M + N + K)
    a = rng.standard_normal((M, K)).astype(np.float16)
    b = rng.standard_normal((K, N)).astype(np.float16)
    d = debug_gemm(a, b, b_mn_major=True)
    ref = _ref(a, b, True)
    err = np.abs(d - ref).max()
    assert err < 2e-3 * np.sqrt(K), (err, d[:2, :4], ref[:2, :4])
