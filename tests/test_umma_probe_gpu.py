"""Pins the tcgen05 shared-memory descriptor conventions and TMEM layouts the conv kernels rely on,
by running single UMMA chains with explicit descriptors and comparing with torch matmul."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _probe(pkg, A, a_box, B, b_box, M, N, nk, a_mn, b_mn, adesc, bdesc, ncols=None):
    L = pkg._lib
    ncols = ncols or max(16, N)
    out = torch.full((128, ncols), float("nan"), device="cuda", dtype=torch.float32)
    ad = (ctypes.c_uint32 * 5)(*adesc)
    bd = (ctypes.c_uint32 * 5)(*bdesc)
    L.check(L.lib().dp_umma_probe(L.ptr(A), A.shape[0], A.shape[1], a_box[0], a_box[1], L.ptr(B), B.shape[0],
                                  B.shape[1], b_box[0], b_box[1], M, N, nk, a_mn, b_mn, ad, bd, L.ptr(out), ncols,
                                  L.stream()))
    torch.cuda.synchronize()
    return out


def _rand(r, c, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randint(-4, 5, (r, c), generator=g).float() / 4).to(torch.bfloat16).cuda()


SW = {128: 2, 64: 4, 32: 6}


@pytest.mark.parametrize("kb", [64, 32, 16])
@pytest.mark.parametrize("N", [16, 64, 128, 256])
def test_k_major_all_swizzles(pkg, kb, N):
    A = _rand(128, kb, 1)
    B = _rand(N, kb, 2)
    rb = kb * 2
    d = [16, 8 * rb, SW[rb], 32, 0]
    out = _probe(pkg, A, (128, kb), B, (N, kb), 128, N, kb // 16, 0, 0, d, d)
    ref = A.float() @ B.float().t()
    assert torch.equal(out[:, :N], ref)


@pytest.mark.parametrize("kb,shift_rows", [(64, 8), (64, 16), (64, 32), (32, 8), (32, 16), (16, 8)])
def test_k_major_row_shifted_start(pkg, kb, shift_rows):
    """the vertical-tap trick: the A descriptor may start any multiple of 8 rows into a taller TMA box."""
    A = _rand(192, kb, 3)
    B = _rand(64, kb, 4)
    rb = kb * 2
    out = _probe(pkg, A, (192, kb), B, (64, kb), 128, 64, kb // 16, 0, 0,
                 [16, 8 * rb, SW[rb], 32, shift_rows * rb], [16, 8 * rb, SW[rb], 32, 0])
    ref = A[shift_rows:shift_rows + 128].float() @ B.float().t()
    assert torch.equal(out[:, :64], ref)


def test_mn_major_a_and_b(pkg):
    """wgrad operands: both matrices stored [K rows][MN contiguous] in 64-wide 128B-swizzled boxes."""
    K, M, N = 64, 128, 64
    At = _rand(K, M, 5)        # [K][M]
    Bt = _rand(K, N, 6)        # [K][N]
    box_bytes = K * 128
    out = _probe(pkg, At, (K, 64), Bt, (K, 64), M, N, K // 16, 1, 1,
                 [box_bytes, 1024, 2, 2048, 0], [box_bytes, 1024, 2, 2048, 0])
    ref = At.float().t() @ Bt.float()
    assert torch.equal(out[:, :N], ref)


def test_mn_major_b_multi_chunk_lbo(pkg):
    """N = 192 as three 64-wide chunks whose smem distance is the descriptor's LBO."""
    K, M = 32, 128
    At = _rand(K, M, 7)
    Bt = _rand(K, 192, 8)
    out = _probe(pkg, At, (K, 64), Bt, (K, 64), M, 192, K // 16, 1, 1,
                 [K * 128, 1024, 2, 2048, 0], [K * 128, 1024, 2, 2048, 0], ncols=192)
    ref = At.float().t() @ Bt.float()
    assert torch.equal(out[:, :192], ref)


def test_m64_tmem_layout(pkg):
    """M=64 accumulators: report where the 64 rows land in the 128 TMEM lanes."""
    A = _rand(64, 64, 9)
    B = _rand(64, 64, 10)
    d = [16, 1024, 2, 32, 0]
    out = _probe(pkg, A, (64, 64), B, (64, 64), 64, 64, 4, 0, 0, d, d)
    ref = A.float() @ B.float().t()
    lanes = []
    for i in range(64):
        hit = [l for l in range(128) if torch.equal(out[l, :64], ref[i])]
        lanes.append(hit[0] if hit else -1)
    print("M=64 row->lane map:", lanes)
    expect = [(i // 16) * 32 + (i % 16) for i in range(64)]
    assert lanes == expect or lanes == list(range(64)), lanes


@pytest.mark.parametrize("mc", [32, 16])
def test_mn_major_narrow_swizzles_and_aliasing(pkg, mc):
    """narrow layers (16/32 channels): MN-major boxes with 64B/32B rows; M=64 built from aliased chunks (LBO=0)."""
    K, N = 64, mc
    At = _rand(K, mc, 12)       # [K][mc]  -> M=64 made of 64/mc aliases of the same chunk
    Bt = _rand(K, N, 13)
    rb = mc * 2
    out = _probe(pkg, At, (K, mc), Bt, (K, N), 64, N, K // 16, 1, 1,
                 [0, 8 * rb, SW[rb], 16 * rb, 0], [0, 8 * rb, SW[rb], 16 * rb, 0], ncols=max(16, N))
    ref = At.float().t() @ Bt.float()       # [mc][N]
    lanes = [(i // 16) * 32 + (i % 16) for i in range(mc)]
    assert torch.equal(out[lanes, :N], ref)


def test_mn_major_b_row_shift(pkg):
    """wgrad vertical taps: the X halo box is read r*tw pixel-rows further down (multiple of 8 rows)."""
    K, M, N = 128, 128, 64
    At = _rand(K, M, 14)
    Xt = _rand(K + 32, N, 15)
    out = _probe(pkg, At, (K, 64), Xt, (K + 32, 64), M, N, K // 16, 1, 1,
                 [K * 128, 1024, 2, 2048, 0], [0, 1024, 2, 2048, 16 * 128])
    ref = At.float().t() @ Xt[16:16 + K].float()
    assert torch.equal(out[:, :N], ref)
