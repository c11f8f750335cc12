"""CPU: host logic of the wide (chfak 2..5) path and the layout arithmetic its TMA / tcgen05 kernels rely on.

No GPU needed: (1) the chunk-planar converters, the coverage predicate and the C-ABI's argument checks (every `cgs_wide_*` entry point
must refuse bad arguments with CGS_EINVAL and a message BEFORE touching a device); (2) a numpy emulation of csrc/wide_tc.cu's address
arithmetic - the TMA box with zero-filled out-of-bounds elements, the UMMA no-swizzle K-major descriptor (start address, LBO, SBO)
shifted per filter tap for the forward / input-gradient GEMM, and the MN-major descriptors of the weight-gradient GEMM with its three
kx-shifted copies and the ones plane - against plain convolutions.  The GPU tests (tests/test_gpu_wide.py) prove the kernels; this
file pins the layout SPEC they were written to, the way tools/hg_emulate.py does for the mma.sync kernels."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F


def test_planar_round_trip_and_supported():
    from cgs_b200 import wide, ops
    from cgs_b200.nets import NewCritic
    x = torch.randn(3, 40, 8, 16)
    p = wide.to_planar(x)
    assert tuple(p.shape) == (3, 5, 8, 16, 8) and p.dtype == torch.bfloat16
    assert torch.equal(wide.from_planar(p), x.to(torch.bfloat16).float())
    # element (n, c, y, x) sits at [n, c // 8, y, x, c % 8]
    assert p[1, 3, 2, 5, 6] == x[1, 3 * 8 + 6, 2, 5].to(torch.bfloat16)
    ops.set_precision("tf32")
    try:
        assert [wide.supported(NewCritic(chfak=k)) for k in (1, 2, 3, 5, 6)] == [False, True, True, True, False]
    finally:
        ops.set_precision("fp32")
    assert not wide.supported(NewCritic(chfak=5))                     # exact-fp32 mode never takes the tensor-core path


def test_wide_entry_points_refuse_bad_arguments_without_a_device():
    from cgs_b200 import _lib
    L = _lib.lib()
    EINVAL = -1
    one = C.c_void_p(16)                                             # a non-NULL, 16-byte aligned dummy: the checks come first
    cases = {
        "conv: NULL input": lambda: L.cgs_wide_conv3x3(None, 2, 8, 8, 40, one, None, None, 40, 0, 0, one, None, None, None, None, None),
        "conv: channels not a multiple of 8": lambda: L.cgs_wide_conv3x3(one, 2, 8, 8, 12, one, None, None, 40, 0, 0, one, None, None, None, None, None),
        "conv: W not a multiple of 8": lambda: L.cgs_wide_conv3x3(one, 2, 8, 12, 40, one, None, None, 40, 0, 0, one, None, None, None, None, None),
        "conv: pooling epilogue without idx_out": lambda: L.cgs_wide_conv3x3(one, 2, 8, 8, 40, one, None, None, 40, 0, 1, one, None, None, None, None, None),
        "conv: unpool epilogue without idx_in": lambda: L.cgs_wide_conv3x3(one, 2, 8, 8, 40, one, None, None, 40, 1, 2, one, None, None, None, None, None),
        "wgrad: Cin > 40": lambda: L.cgs_wide_wgrad3x3(one, one, 2, 8, 8, 48, 40, one, one, one, 1 << 30, None),
        "wgrad: NULL workspace": lambda: L.cgs_wide_wgrad3x3(one, one, 2, 8, 8, 40, 40, one, one, None, 0, None),
        "conv0: C0 > 40": lambda: L.cgs_wide_conv0_fwd(one, 2, 0, None, one, one, 48, one, one, None),
        "gemm: unaligned leading dimension": lambda: L.cgs_wide_gemm(one, 1, 30, one, 1, 32, one, 32, 4, 32, 30, None, None, 0, 0, 1, None, None, None),
        "gemm: split-K without workspace": lambda: L.cgs_wide_gemm(one, 1, 32, one, 1, 32, one, 32, 4, 32, 32, None, None, 0, 0, 4, None, None, None),
        "pack: no jobs": lambda: L.cgs_wide_pack(None, 0, None),
        "colsums: too many jobs": lambda: L.cgs_wide_colsums(one, 6, 4, None),
        "critic_train_bf16: NULL partials": lambda: L.cgs_critic_train_bf16(one, one, 4, 0, None, None, None, None, 0.0, 0, None, C.byref(_lib.CriticWeights()),
                                                                            None, None, 1.0, 0, one, one, None),
    }
    for what, call in cases.items():
        rc = call()
        assert rc == EINVAL, (what, rc)
        assert L.cgs_last_error(), what
    assert L.cgs_wide_packed_bytes(40, 40) == 9 * 6 * 48 * 16 and L.cgs_wide_packed_bytes(80, 40) == 9 * 10 * 48 * 16


# ---------------------------------------------------------------------------------------------------------------------
# numpy model of the shared-memory side of csrc/wide_tc.cu
def _bf(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).float().numpy()


def tma_box(act, n, plane0, planes, y0, x0, bh, bw):
    """Box (8 bw, bh, planes, 1) of the 4-D view (8 W, H, C/8, B) of a chunk-planar activation [B][C/8][H][W][8] at pixel coordinates
    (x0, y0): lands as [plane][bh][bw] slots of 8 values (16 bytes); out-of-bounds elements are zero (the convolution's padding)."""
    B, CP, H, W, _ = act.shape
    out = np.zeros((planes, bh, bw, 8), np.float32)
    for p in range(planes):
        for r in range(bh):
            for c in range(bw):
                y, x = y0 + r, x0 + c
                if 0 <= y < H and 0 <= x < W and plane0 + p < CP:
                    out[p, r, c] = act[n, plane0 + p, y, x]
    return out


def umma_operand(smem_slots, start, lbo, sbo, rows, k_major):
    """Read a rows x 16 bf16 operand through a no-swizzle UMMA descriptor (all quantities in 16-byte slots).
    K-major  ((8, m), (8, 2)) : ((1 slot, SBO), (elements, LBO)): row r, k -> slot start + (r % 8) + (r // 8) * SBO + (k // 8) * LBO, element k % 8.
    MN-major ((8, m), (8, 2)) : ((elements, SBO), (1 slot, LBO)): row r, k -> slot start + (k % 8) + (k // 8) * LBO + (r // 8) * SBO, element r % 8."""
    out = np.zeros((rows, 16), np.float32)
    for r in range(rows):
        for k in range(16):
            if k_major:
                out[r, k] = smem_slots[start + (r % 8) + (r // 8) * sbo + (k // 8) * lbo, k % 8]
            else:
                out[r, k] = smem_slots[start + (k % 8) + (k // 8) * lbo + (r // 8) * sbo, r % 8]
    return out


@pytest.mark.parametrize("Cin,Cout,H,W,transposed", [(40, 16, 16, 8, False), (16, 24, 8, 8, False), (24, 16, 16, 16, True)])
def test_forward_gemm_is_nine_shifted_descriptors_over_one_haloed_tile(Cin, Cout, H, W, transposed):
    """wide_conv_kernel: A = the haloed 18 x 10 tile ([plane][18][10] slots: SBO = 10, LBO = 180), tap (ky, kx) = the same buffer with the
    start address moved by ky * 10 + kx slots; B = the packed filters [(tap * KP + plane) * NP + co] (SBO = 8 rows, LBO = NP)."""
    rng = np.random.default_rng(0)
    B = 2
    x = _bf(rng.standard_normal((B, Cin, H, W)).astype(np.float32))
    w = rng.standard_normal((Cin, Cout, 3, 3) if transposed else (Cout, Cin, 3, 3)).astype(np.float32) * 0.2
    ref = (F.conv_transpose2d if transposed else F.conv2d)(torch.from_numpy(x), torch.from_numpy(_bf(w)), padding=1).numpy()
    act = x.reshape(B, Cin // 8, 8, H, W).transpose(0, 1, 3, 4, 2)
    CPi, KP, NP = Cin // 8, (Cin // 8 + 1) & ~1, (Cout + 15) & ~15
    packed = np.zeros((9 * KP * NP, 8), np.float32)                  # pack_row() of csrc/wide_tc.cu
    for r in range(9 * KP * NP):
        co, q, t = r % NP, (r // NP) % KP, r // NP // KP
        for c8 in range(8):
            ci = q * 8 + c8
            if ci < Cin and co < Cout:
                packed[r, c8] = w[ci, co].reshape(9)[8 - t] if transposed else w[co, ci].reshape(9)[t]
    packed = _bf(packed)
    for n in range(B):
        for ty in range((H + 15) // 16):
            for tx in range(W // 8):
                tile = np.zeros((KP, 18, 10, 8), np.float32)        # the pad plane (Cin / 8 odd) is zeroed by the kernel
                tile[:CPi] = tma_box(act, n, 0, CPi, ty * 16 - 1, tx * 8 - 1, 18, 10)
                slots = tile.reshape(-1, 8)
                D = np.zeros((128, NP), np.float64)
                for tap in range(9):
                    for kp in range(0, KP, 2):
                        A = umma_operand(slots, (tap // 3) * 10 + tap % 3 + kp * 180, 180, 10, 128, True)
                        Bm = umma_operand(packed, (tap * KP + kp) * NP, NP, 8, NP, True)
                        D += A.astype(np.float64) @ Bm.astype(np.float64).T
                for m in range(128):                                 # accumulator lane m = pixel (m // 8, m % 8) of the tile
                    y, xx = ty * 16 + m // 8, tx * 8 + m % 8
                    if y < H:
                        assert np.allclose(D[m, :Cout], ref[n, :, y, xx], rtol=1e-5, atol=1e-5), (n, ty, tx, m)


def test_weight_gradient_gemm_is_mn_major_with_three_kx_copies_and_a_ones_plane():
    """wide_wgrad_kernel: A = planes [kx][ci / 8] (+ plane 15 = ones) of 18 x 8 slots (three TMA boxes shifted by kx), B = the 16 x 8 dY
    tile; both MN-major (LBO = one tile row = 8 slots, SBO = one plane); per ky and row pair r2 one MMA with K = 16 pixels; D_ky[(kx, ci)][co]."""
    rng = np.random.default_rng(1)
    B, Cin, Cout, H, W = 2, 24, 16, 16, 16
    x = _bf(rng.standard_normal((B, Cin, H, W)).astype(np.float32))
    dy = _bf(rng.standard_normal((B, Cout, H, W)).astype(np.float32))
    dw_ref = torch.nn.grad.conv2d_weight(torch.from_numpy(x).double(), (Cout, Cin, 3, 3), torch.from_numpy(dy).double(), padding=1).numpy()
    db_ref = dy.astype(np.float64).sum((0, 2, 3))
    ax = x.reshape(B, Cin // 8, 8, H, W).transpose(0, 1, 3, 4, 2)
    ady = dy.reshape(B, Cout // 8, 8, H, W).transpose(0, 1, 3, 4, 2)
    CPi, CPo = Cin // 8, Cout // 8
    D = np.zeros((3, 128, Cout), np.float64)
    for n in range(B):
        for ty in range(H // 16):
            for tx in range(W // 8):
                a = np.zeros((16, 18, 8, 8), np.float32)
                for kx in range(3):
                    a[kx * CPi:(kx + 1) * CPi] = tma_box(ax, n, 0, CPi, ty * 16 - 1, tx * 8 - 1 + kx, 18, 8)
                a[15] = 1.0
                b = tma_box(ady, n, 0, CPo, ty * 16, tx * 8, 16, 8)
                sa, sb = a.reshape(-1, 8), b.reshape(-1, 8)
                for ky in range(3):
                    for r2 in range(8):
                        A = umma_operand(sa, (2 * r2 + ky) * 8, 8, 18 * 8, 128, False)
                        Bm = umma_operand(sb, 2 * r2 * 8, 8, 16 * 8, Cout, False)
                        D[ky] += A.astype(np.float64) @ Bm.astype(np.float64).T
    for ky in range(3):
        for kx in range(3):
            for ci in range(Cin):
                assert np.allclose(D[ky, kx * Cin + ci], dw_ref[:, ci, ky, kx], rtol=1e-6, atol=1e-6)
    assert np.allclose(D[1, 120], db_ref, rtol=1e-6, atol=1e-6)      # the ones plane: every row of it holds the bias gradient


def test_pooling_epilogue_channel_split_matches_max_pool_with_indices():
    """The conv kernel's pooling epilogue: the 4 lanes of a 2x2 window (l, l^1, l^8, l^9) exchange so that each finishes 2 of the 8
    channels of a group (x exchange: even x keeps channels 0-3; y exchange: even y keeps the first two of those); first maximum in
    row-major order wins, 4 = pooled value not > 0.  Emulated lane by lane against F.max_pool2d(F.relu(.), 2, return_indices=True)."""
    rng = np.random.default_rng(2)
    vals = rng.standard_normal((32, 8)).astype(np.float32)            # one warp = 4 tile rows x 8 pixels, 8 channels each
    vals[5] = vals[4]                                                # ties inside a window
    vals[12:14, 3] = -1.0; vals[4:6, 3] = -2.0                       # a window that is dead in channel 3
    out = np.zeros((2, 4, 8), np.float32); idx = np.zeros((2, 4, 8), np.int64)
    mx = np.zeros((32, 4), np.float32); rb = np.zeros((32, 4), np.int64)
    for lane in range(32):
        odd_x = lane & 1
        for i in range(4):
            recv = vals[lane ^ 1][4 + i] if odd_x else vals[lane ^ 1][i]      # the partner sends the half this lane keeps
            mine = vals[lane][4 + i] if odd_x else vals[lane][i]
            left, right = (recv, mine) if odd_x else (mine, recv)
            rb[lane, i] = int(right > left); mx[lane, i] = right if right > left else left
    for lane in range(32):
        odd_x, odd_y = lane & 1, (lane >> 3) & 1
        other = lane ^ 8
        for k in range(2):
            mine = mx[lane][2 + k] if odd_y else mx[lane][k]
            recv = mx[other][2 + k] if odd_y else mx[other][k]
            top, bot = (recv, mine) if odd_y else (mine, recv)
            tb = rb[other][2 + k] if odd_y else rb[lane][k]
            bb = rb[lane][2 + k] if odd_y else rb[other][k]
            v, am = (bot, 2 + bb) if bot > top else (top, tb)
            if not v > 0:
                v, am = 0.0, 4
            c = (4 if odd_x else 0) + (2 if odd_y else 0) + k
            out[(lane >> 4), (lane & 7) >> 1, c] = v; idx[(lane >> 4), (lane & 7) >> 1, c] = am
    t = torch.from_numpy(vals.reshape(4, 8, 8)).permute(2, 0, 1)[None]    # [1, C, 4 rows, 8 px]
    pr, pi = F.max_pool2d(F.relu(t), 2, return_indices=True)
    pos = ((pi // 8) % 2) * 2 + (pi % 8) % 2
    pr, pos = pr[0].permute(1, 2, 0).numpy(), pos[0].permute(1, 2, 0).numpy()
    assert np.array_equal(out, pr)
    live = pr > 0
    assert np.array_equal(idx[live], pos[live]) and (idx[~live] == 4).all()
