"""CPU: the fragment addressing of the batched decoder's skinny GEMM (csrc/decoder_batch.cu `bd_gemm_kernel`), emulated in
numpy with the PTX fragment layout of mma.m16n8k16 (row-major A 16x16, column-major B 16x8, C 16x8):

    a0 = A[g][2t, 2t+1]   a1 = A[g+8][2t, 2t+1]   a2 = A[g][2t+8, 2t+9]   a3 = A[g+8][2t+8, 2t+9]
    b0 = B[2t, 2t+1][g]   b1 = B[2t+8, 2t+9][g]
    c0 = C[g][2t]   c1 = C[g][2t+1]   c2 = C[g+8][2t]   c3 = C[g+8][2t+1]          (g = lane / 4, t = lane % 4)

The kernel loads, per 32-wide K block, the 8 halves [8t, 8t+8) of weight rows g / g+8 and of x row g as ONE 16-byte vector each
and feeds register pairs (x, y) and (z, w) to two MMAs.  The MMA therefore sees K in a permuted order - the same permutation on
both operands - and this test checks that every product is still formed exactly once for all warp tilings the kernel is
instantiated with, including the clamped rows past N, the K split across warps and the shared-memory fold's index math."""
import numpy as np
import pytest


def _emulate(WR, WK, NT, N, K, B, rng):
    W = rng.standard_normal((N, K)).astype(np.float32)
    X = np.zeros((8 * NT, K), np.float32)
    X[:B] = rng.standard_normal((B, K))
    out = np.full((B, N), np.nan, np.float32)
    lanes = [(lane >> 2, lane & 3) for lane in range(32)]
    for cta in range((N + 16 * WR - 1) // (16 * WR)):
        red = np.zeros((WR, WK, 8 * NT, 17), np.float32)
        for warp in range(8):
            wr, wk = warp // WK, warp % WK
            row_base = (cta * WR + wr) * 16
            nblk = K >> 5
            blk0, blk1 = wk * nblk // WK, (wk + 1) * nblk // WK
            C = np.zeros((NT, 16, 8), np.float32)
            for blk in range(blk0, blk1):
                for half in range(2):                      # the two MMAs of a block: vector components (x, y) then (z, w)
                    A = np.zeros((16, 16), np.float32)
                    Bm = np.zeros((NT, 16, 8), np.float32)
                    o = 4 * half
                    for g, t in lanes:
                        ra, rb = min(row_base + g, N - 1), min(row_base + g + 8, N - 1)
                        a = W[ra, blk * 32 + t * 8: blk * 32 + t * 8 + 8]
                        c = W[rb, blk * 32 + t * 8: blk * 32 + t * 8 + 8]
                        A[g, 2 * t:2 * t + 2] = a[o:o + 2]; A[g + 8, 2 * t:2 * t + 2] = c[o:o + 2]
                        A[g, 2 * t + 8:2 * t + 10] = a[o + 2:o + 4]; A[g + 8, 2 * t + 8:2 * t + 10] = c[o + 2:o + 4]
                        for nt in range(NT):
                            xv = X[nt * 8 + g, blk * 32 + t * 8: blk * 32 + t * 8 + 8]
                            Bm[nt, 2 * t:2 * t + 2, g] = xv[o:o + 2]; Bm[nt, 2 * t + 8:2 * t + 10, g] = xv[o + 2:o + 4]
                    for nt in range(NT):
                        C[nt] += A @ Bm[nt]
            for g, t in lanes:
                for nt in range(NT):
                    red[wr, wk, nt * 8 + 2 * t, g] = C[nt, g, 2 * t]; red[wr, wk, nt * 8 + 2 * t + 1, g] = C[nt, g, 2 * t + 1]
                    red[wr, wk, nt * 8 + 2 * t, g + 8] = C[nt, g + 8, 2 * t]; red[wr, wk, nt * 8 + 2 * t + 1, g + 8] = C[nt, g + 8, 2 * t + 1]
        total = WR * 16 * 8 * NT
        for tid in range(256):
            for k in range((total + 255) // 256):
                idx = tid + k * 256
                if idx >= total:
                    break
                r, b, w2 = idx & 15, (idx >> 4) % (8 * NT), idx // (16 * 8 * NT)
                row = (cta * WR + w2) * 16 + r
                if row < N and b < B:
                    assert np.isnan(out[b, row])           # every output is written by exactly one thread
                    out[b, row] = red[w2, :, b, r].sum()
    return out, X[:B] @ W.T


@pytest.mark.parametrize("WR,WK,NT,N,K,B", [
    (1, 8, 1, 100, 256, 5),      # N = d tilings, one n-tile, ragged N
    (1, 8, 4, 64, 1024, 32),     # full batch
    (2, 4, 2, 200, 384, 11),     # QKV / FC1 tiling, K blocks that do not divide evenly among the K splits (12 over 4)
    (8, 1, 4, 300, 128, 29),     # LM-head tiling: no K split, N not a multiple of 128
    (1, 8, 2, 48, 384, 16),      # 12 K blocks over 8 splits: some warps get one block, some two
])
def test_permuted_k_fragments_form_every_product_once(WR, WK, NT, N, K, B):
    out, ref = _emulate(WR, WK, NT, N, K, B, np.random.default_rng(N + K))
    assert not np.isnan(out).any()
    np.testing.assert_allclose(out, ref, rtol=0, atol=2e-4)


def _simulate_x_pipeline(n_blocks, U=5, XD=3):
    """Emulates the register pipeline of bd_gemm_kernel's main loop (decoder_batch.cu): XD x-slots, U weight registers per round,
    refills behind the products.  Returns the list of (x block, weight block) pairs that were multiplied."""
    blk0, blk1 = 0, n_blocks
    a = [min(blk0 + u, blk1 - 1) for u in range(U)]                 # weight block held by register u
    xq = [min(blk0 + u, blk1 - 1) for u in range(XD)]               # x block held by slot u
    done = []
    blk = blk0
    while blk < blk1:
        more = blk + U < blk1
        for u in range(U):
            on = blk + u < blk1
            if on:
                done.append((xq[u % XD], a[u]))
            bx = blk + u + XD if u + XD < U else blk + U + (u % XD)
            if bx < blk1:
                xq[u % XD] = bx
            if more:
                a[u] = min(blk + U + u, blk1 - 1)
        blk += U
    return done


def test_batched_gemm_x_pipeline_schedule():
    """Every K block of a warp is multiplied exactly once, x block against the weight block of the same index - for one round (5 blocks:
    d = 1280 split 8 ways), whole rounds (10, 20: QKV / FC1 / FC2 of large-v3) and short tails (1, 2, 3, 6, 7, 12: the small models)."""
    for n in (1, 2, 3, 4, 5, 6, 7, 9, 10, 11, 12, 15, 20, 23):
        done = _simulate_x_pipeline(n)
        assert done == [(b, b) for b in range(n)], (n, done)


def test_decode_kernel_diagonal_reduction_lanes():
    """decoder_mega.cu, gemv_phase: one m16n8k16 tile = {2 weight rows} x {8 K slices}; the wanted products are the diagonal
    C[8 * row + slice][slice].  In the C fragment lane (g, t) holds C[g][2t], C[g][2t+1], C[g+8][2t], C[g+8][2t+1], so slice s of
    both rows sits on lane 4s + (s >> 1) (element s & 1).  The kernel folds the 8 slices with xor-shuffles 4, 9, 18 and lets
    diagonal lane number e finish row e & 1 of chunk e >> 1; the odd row's value reaches the even row's lane by shfl_down 4."""
    lanes = []
    for s in range(8):
        cand = [lane for lane in range(32) if (lane >> 2) == s and (lane & 3) == (s >> 1)]      # g == s and the column pair holding column s
        assert len(cand) == 1
        lanes.append(cand[0])
    assert lanes == [4 * s + (s >> 1) for s in range(8)] == [0, 4, 9, 13, 18, 22, 27, 31]
    # the kernel's test for "diagonal lane" and its index
    for lane in range(32):
        diag = (lane & 3) == (lane >> 3)
        assert diag == (lane in lanes)
        if diag:
            assert lanes[lane >> 2] == lane
    # three xor levels reach every diagonal lane from every diagonal lane, and never leave the set
    vals = np.zeros(32)
    vals[lanes] = np.arange(1, 9) * 1.5
    v = vals.copy()
    for mask in (4, 9, 18):
        v = v + v[np.arange(32) ^ mask]
        assert all(((lane ^ mask) in lanes) for lane in lanes)
    assert np.allclose(v[lanes], vals.sum())
    # finishing lanes: e = 2 * chunk + row; the pair's odd row is 4 lanes further down
    for e in range(0, 8, 2):
        assert lanes[e + 1] - lanes[e] == 4
