"""tcgen05/TMA GEMM against numpy (f16 inputs, f32 accumulate): exact products, so tolerance only
covers summation order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ref(a, b, mn):
    a32, b32 = a.astype(np.float32), b.astype(np.float32)
    return a32 @ (b32 if mn else b32.T)


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 256), (300, 200, 192), (1500, 1280, 1280),
                                   (1500, 1500, 64), (257, 130, 240), (3000, 384, 240),
                                   (1500, 3840, 1280), (4000, 5120, 256), (129, 257, 64)])
def test_gemm_k_major(M, N, K):
    from speaksense_b200.asr import debug_gemm
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    a = rng.standard_normal((M, K)).astype(np.float16)
    b = rng.standard_normal((N, K)).astype(np.float16)
    d = debug_gemm(a, b)
    ref = _ref(a, b, False)
    err = np.abs(d - ref).max()
    assert err < 2e-3 * np.sqrt(K), (err, d[:2, :4], ref[:2, :4])


def test_gemm_b_mn_major_is_attention_only():
    from speaksense_b200 import NativeError
    from speaksense_b200.asr import debug_gemm
    a = np.zeros((128, 64), np.float16)
    with pytest.raises(NativeError):
        debug_gemm(a, np.zeros((64, 64), np.float16), b_mn_major=True)
