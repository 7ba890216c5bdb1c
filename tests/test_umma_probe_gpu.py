"""Pins the tcgen05 shared-memory descriptor conventions and TMEM layouts the conv kernels rely on,
by running single UMMA chains with explicit descriptors and comparing with torch matmul."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


_PROBE = None


def _probe_lib():
    """tests/probe/libdepth_b200_probe.so (built by build.py next to the product library, not part of it)"""
    global _PROBE
    if _PROBE is None:
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "probe", "libdepth_b200_probe.so")
        lib = ctypes.CDLL(path)
        P, I = ctypes.c_void_p, ctypes.c_int
        lib.dp_umma_probe.restype = I
        lib.dp_umma_probe.argtypes = [P, I, I, I, I, P, I, I, I, I, I, I, I, I, I, ctypes.POINTER(ctypes.c_uint32),
                                      ctypes.POINTER(ctypes.c_uint32), P, I, P]
        _PROBE = lib
    return _PROBE


def _probe(pkg, A, a_box, B, b_box, M, N, nk, a_mn, b_mn, adesc, bdesc, ncols=None):
    L = pkg._lib
    ncols = ncols or max(16, N)
    out = torch.full((128, ncols), float("nan"), device="cuda", dtype=torch.float32)
    ad = (ctypes.c_uint32 * 6)(*(list(adesc) + [0] * (6 - len(adesc))))
    bd = (ctypes.c_uint32 * 5)(*bdesc)
    rc = (_probe_lib().dp_umma_probe(L.ptr(A), A.shape[0], A.shape[1], a_box[0], a_box[1], L.ptr(B), B.shape[0],
                                  B.shape[1], b_box[0], b_box[1], M, N, nk, a_mn, b_mn, ad, bd, L.ptr(out), ncols,
                                  L.stream()))
    assert rc == 0, f"dp_umma_probe returned {rc}"
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


@pytest.mark.parametrize("kb", [64, 32, 16])
@pytest.mark.parametrize("shift", [1, 2, 11, 21, 22])
def test_k_major_halo_addressing(pkg, kb, shift):
    """single-halo-load conv: the swizzle is a function of the absolute shared-memory address, so an A descriptor may
    start ANY whole row into a swizzled TMA box and step between its 8-row groups with a stride (SBO) that is the halo
    row pitch (tw+2 = 10 rows), not a multiple of the swizzle atom.  One box then serves all nine 3x3 taps."""
    A = _rand(256, kb, 41)
    B = _rand(64, kb, 42)
    rb = kb * 2
    out = _probe(pkg, A, (256, kb), B, (64, kb), 128, 64, kb // 16, 0, 0,
                 [16, 10 * rb, SW[rb], 32, shift * rb, 0], [16, 8 * rb, SW[rb], 32, 0])
    rows = torch.cat([A[g * 10 + shift: g * 10 + shift + 8] for g in range(16)])
    ref = rows.float() @ B.float().t()
    assert torch.equal(out[:, :64], ref)


@pytest.mark.parametrize("nc", [64, 32, 16])
@pytest.mark.parametrize("shift", [0, 1, 12])
def test_mn_major_halo_addressing(pkg, nc, shift):
    """wgrad from one halo box: MN-major B operand whose K (pixel) groups of 8 rows are a halo pitch (10 rows) apart and
    whose N chunks (vertical taps) are LBO = one halo row pitch apart; start shifted by whole rows."""
    K, M = 64, 128                                  # 64 pixels = 8 patch rows of 8
    At = _rand(K, M, 43)
    Xt = _rand(140, nc, 44)                         # halo box rows
    rbB = nc * 2
    out = _probe(pkg, At, (K, 64), Xt, (140, nc), M, 3 * nc, K // 16, 1, 1,
                 [K * 128, 1024, 2, 2048, 0], [10 * rbB, 10 * rbB, SW[rbB], 20 * rbB, shift * rbB], ncols=max(16, 3 * nc))
    ref = torch.zeros(M, 3 * nc, device="cuda")
    for r in range(3):
        rows = torch.cat([Xt[(g + r) * 10 + shift: (g + r) * 10 + shift + 8] for g in range(8)])   # [64][nc]
        ref[:, r * nc:(r + 1) * nc] = At.float().t() @ rows.float()
    assert torch.equal(out[:, :3 * nc], ref)


def test_explore_unaligned_row_shift(pkg):
    """exploration (prints, asserts nothing about the outcome): can the A descriptor start a non-multiple-of-8 rows
    into a 128B-swizzled box (horizontal conv tap = +1 pixel)?  Tries base_offset = 0 and = shift."""
    A = _rand(256, 64, 41)
    B = _rand(64, 64, 42)
    res = {}
    for shift in (1, 2, 3, 9):
        for boff in (0, shift & 7):
            out = _probe(pkg, A, (256, 64), B, (64, 64), 128, 64, 4, 0, 0,
                         [16, 1024, 2, 32, shift * 128, boff], [16, 1024, 2, 32, 0])
            ref = A[shift:shift + 128].float() @ B.float().t()
            res[(shift, boff)] = bool(torch.equal(out[:, :64], ref))
    # 8-pixel-wide patch inside a 16-pixel-pitch halo box: group stride (SBO) = 2048, start shifted by s rows
    for shift in (1, 2):
        for boff in (0, shift):
            out = _probe(pkg, A, (256, 64), B, (64, 64), 128, 64, 4, 0, 0,
                         [16, 2048, 2, 32, shift * 128, boff], [16, 1024, 2, 32, 0])
            rows = torch.cat([A[g * 16 + shift: g * 16 + shift + 8] for g in range(16)])[:128]
            ref = rows.float() @ B.float().t()
            res[("pitch16", shift, boff)] = bool(torch.equal(out[:, :64], ref))
    for shift in (0, 1, 2):
        out = _probe(pkg, A, (256, 64), B, (64, 64), 128, 64, 4, 0, 0,
                     [16, 1280, 2, 32, shift * 128, 0], [16, 1024, 2, 32, 0])
        rows = torch.cat([A[g * 10 + shift: g * 10 + shift + 8] for g in range(16)])[:128]
        ref = rows.float() @ B.float().t()
        res[("pitch10", shift)] = bool(torch.equal(out[:, :64], ref))
    print("UNALIGNED-SHIFT EXPLORATION:", res)
