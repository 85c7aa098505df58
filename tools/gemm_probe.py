"""GEMM shapes of the large-v3 encoder through ss_debug_gemm (for `ncu --metrics gpu__time_duration.sum -k regex:gemm_tcgen05`)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from speaksense_b200.asr import debug_gemm
rng = np.random.default_rng(0)
for M, N, K in [(1500, 5120, 1280), (1500, 1280, 5120), (1500, 3840, 1280), (1500, 1280, 1280), (48000, 5120, 1280)]:
    a = rng.standard_normal((M, K)).astype(np.float16); b = rng.standard_normal((N, K)).astype(np.float16)
    debug_gemm(a, b); debug_gemm(a, b)
    print(M, N, K, flush=True)
