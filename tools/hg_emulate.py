"""CPU emulation of the warp-level data paths of csrc/hg_*.cu (authoring aid: there is no GPU in the authoring container).

A warp is emulated lane by lane with the PTX-defined fragment layouts of ldmatrix(.trans) and mma.sync.m16n8k16, the
address formulas are transcribed from the kernels, and every emulated phase is compared with a plain numpy convolution /
weight gradient.  What this checks: operand layouts, tap pairing, the pack_wk() weight-fragment tables, the sliding
row reuse, the .trans weight-gradient operands and the hand-gathered masker.2 fragments.  Run: python tools/hg_emulate.py"""
import numpy as np

P1, P2, PX, DLP = 34, 18, 66, 68
rng = np.random.default_rng(0)


# ------------------------------------------------------------------ PTX fragment semantics (values kept as float64)
def ldsm(read_row, addrs, n, trans):
    """ldmatrix.x{n}[.trans] .b16: addrs[lane] = row address supplied by each lane; read_row(addr) -> 8 values (one 16-byte
    row).  Returns regs[lane][j] = (lo, hi)."""
    regs = [[None] * n for _ in range(32)]
    for j in range(n):
        M = np.stack([read_row(addrs[8 * j + r]) for r in range(8)])          # [row][col]
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            regs[lane][j] = (M[2 * t][g], M[2 * t + 1][g]) if trans else (M[g][2 * t], M[g][2 * t + 1])
    return regs


def mma(C, A, B):
    """C[lane][4] += A x B; A[lane] = 4 regs (lo, hi), B[lane] = 2 regs."""
    a = np.zeros((16, 16)); b = np.zeros((16, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for j, (row, k0) in enumerate(((g, 2 * t), (g + 8, 2 * t), (g, 2 * t + 8), (g + 8, 2 * t + 8))):
            a[row, k0], a[row, k0 + 1] = A[lane][j]
        for j, k0 in enumerate((2 * t, 2 * t + 8)):
            b[k0, g], b[k0 + 1, g] = B[lane][j]
    c = a @ b
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        C[lane] += np.array([c[g, 2 * t], c[g, 2 * t + 1], c[g + 8, 2 * t], c[g + 8, 2 * t + 1]])


ZERO = (0.0, 0.0)


# ------------------------------------------------------------------ pack_wk (transcribed from hg_forward.cu)
F_C0, F_C1, F_C2, F_C3, F_D2, F_D1, F_D0, F_M0, F_M2, F_D3 = 0, 3, 9, 15, 25, 43, 52, 61, 79, 88
B_M0D, B_D0D, B_D1D, B_M2D, B_D2D, B_D3D = 142, 151, 157, 163, 165, 175
B_C3D, B_C2D, B_C1D, B_C0D, NSTEPS = 211, 220, 226, 232, 238


def pack_wk(p, s, k, n):
    if s < F_C1:
        ky, kx, c = s, k >> 2, k & 3
        return p["w0"][n, c, ky, kx] if (kx < 3 and c < 3) else 0.0
    if s < F_C3:
        w = p["w1"] if s < F_C2 else p["w2"]
        q = s - F_C1 if s < F_C2 else s - F_C2
        ky, h = q >> 1, q & 1
        if h and k >= 8:
            return 0.0
        kx, ci = (2 if h else (k >> 3)), k & 7
        return w[n, ci, ky, kx]
    if s < F_D2:
        q = s - F_C3; tp, nt = q >> 1, q & 1; tap, ci = 2 * tp + (k >> 3), k & 7
        return 0.0 if tap > 8 else p["w3"].reshape(16, 8, 9)[nt * 8 + n, ci, tap]
    if s < F_D1:
        q = s - F_D2; tap, kc = q >> 1, q & 1
        if kc and k >= 8:
            return 0.0
        return p["d2"].reshape(8, 24, 9)[n, kc * 16 + k, tap]
    if s < F_M0:
        w = p["d1"] if s < F_D0 else p["d0"]
        tap = s - F_D1 if s < F_D0 else s - F_D0
        return w.reshape(8, 16, 9)[n, k, tap]
    if s < F_M2:
        q = s - F_M0; nt, j, ky = q & 1, (q >> 1) % 3, (q >> 1) // 3; co = nt * 8 + n
        if j == 0:
            kx, c = k >> 2, k & 3
            return p["m0"][co, c, ky, kx] if (kx < 3 and c < 3) else 0.0
        if j == 2 and k >= 8:
            return 0.0
        kx, ci = ((k >> 3) if j == 1 else 2), 3 + (k & 7)
        return p["m0"][co, ci, ky, kx]
    if s < F_D3:
        return p["m2"].reshape(16, 9)[k, s - F_M2] if n == 0 else 0.0
    if s < B_M0D:
        q = s - F_D3; nt, kc, tap = q & 1, (q >> 1) % 3, (q >> 1) // 3
        return p["d3"].reshape(16, 48, 9)[nt * 8 + n, kc * 16 + k, tap]
    if s < B_D0D:
        return p["m0"].reshape(16, 11, 9)[k, 3 + n, 8 - (s - B_M0D)]
    if s < B_M2D:
        w = p["d0"] if s < B_D1D else p["d1"]
        q = s - B_D0D if s < B_D1D else s - B_D1D
        ky, h = q >> 1, q & 1
        if h and k >= 8:
            return 0.0
        kx, co = (2 if h else (k >> 3)), k & 7
        return w.reshape(8, 16, 9)[co, 8 + n, 8 - (ky * 3 + kx)]
    if s < B_D2D:
        return p["m2"].reshape(16, 9)[(s - B_M2D) * 8 + n, 8 - k] if k < 9 else 0.0
    if s < B_D3D:
        q = s - B_D2D; nt = q & 1; tp = 2 * (q >> 1) + (k >> 3); co = k & 7
        return 0.0 if tp > 8 else p["d2"].reshape(8, 24, 9)[co, 8 + nt * 8 + n, 8 - tp]
    if s < B_C3D:
        q = s - B_D3D; nt, tp = q & 3, q >> 2
        return p["d3"].reshape(16, 48, 9)[k, 16 + nt * 8 + n, 8 - tp]
    if s < B_C2D:
        return p["w3"].reshape(16, 8, 9)[k, n, 8 - (s - B_C3D)]
    w = p["w2"] if s < B_C1D else (p["w1"] if s < B_C0D else p["w0"])
    q = s - (B_C2D if s < B_C1D else (B_C1D if s < B_C0D else B_C0D))
    ky, h = q >> 1, q & 1
    if h and k >= 8:
        return 0.0
    kx, co = (2 if h else (k >> 3)), k & 7
    tap = 8 - (ky * 3 + kx)
    if s < B_C0D:
        return w.reshape(8, 8, 9)[co, n, tap]
    return w.reshape(8, 3, 9)[co, n, tap] if n < 3 else 0.0


def wfrag(p, s):
    """B fragment of step s for every lane: [(b0), (b1)]"""
    out = []
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        out.append([(pack_wk(p, s, 2 * t, g), pack_wk(p, s, 2 * t + 1, g)), (pack_wk(p, s, 2 * t + 8, g), pack_wk(p, s, 2 * t + 9, g))])
    return out


# ------------------------------------------------------------------ planes
def plane(x_hwc):
    """[H][W][8] -> haloed [H+2][W+2][8]; 'address' = (row, col) pixel index"""
    H, W, _ = x_hwc.shape
    p = np.zeros((H + 2, W + 2, 8))
    p[1:-1, 1:-1] = x_hwc
    return p


def pairdup(x_hw3, nrows=66):
    """[64][64][3] -> [66][66][8]: entry e = {haloed pixel e (rgb0), haloed pixel e+1 (rgb0)}"""
    h = np.zeros((66, 67, 4))
    h[1:65, 1:65, :3] = x_hw3
    out = np.zeros((66, 66, 8))
    out[:, :, :4] = h[:, :66]
    out[:, :, 4:] = h[:, 1:67]
    return out


def lanes():
    for lane in range(32):
        yield lane, lane >> 3, lane & 7, lane >> 2, lane & 3


def conv_ref(x_chw, w, pad=1):
    """plain cross-correlation, x [C][H][W], w [O][C][3][3] -> [O][H][W]"""
    C, H, W = x_chw.shape
    xp = np.zeros((C, H + 2, W + 2)); xp[:, 1:-1, 1:-1] = x_chw
    out = np.zeros((w.shape[0], H, W))
    for ky in range(3):
        for kx in range(3):
            out += np.einsum("oc,chw->ohw", w[:, :, ky, kx], xp[:, ky:ky + H, kx:kx + W])
    return out


def slide(R, NK, w, loadA):
    """slide_bf for one channel tile: returns rows[oi][lane][4]"""
    acc = {}
    for i in range(R + 2):
        a = loadA(i)
        for ky in range(3):
            oi = i - ky
            if 0 <= oi < R:
                if ky == 0:
                    acc[oi] = [np.zeros(4) for _ in range(32)]
                for kk in range(NK):
                    mma(acc[oi], a[kk], w[ky][kk])
    return acc


# ------------------------------------------------------------------ checks
def check_conv0(p):
    x = rng.random((64, 64, 3))
    X = pairdup(x)
    ref = conv_ref(x.transpose(2, 0, 1), p["w0"])
    w = [[wfrag(p, F_C0 + ky)] for ky in range(3)]
    x0, r0 = 16, 32
    def loadA(i):
        addrs = [(r0 + i, x0 + (lr + 8 * (lj & 1)) + 2 * (lj >> 1)) for _, lj, lr, _, _ in lanes()]
        return [ldsm(lambda a: X[a[0], a[1]], addrs, 4, False)]
    acc = slide(16, 1, w, loadA)
    err = 0.0
    for oi, rows in acc.items():
        for lane, _, _, g, t in lanes():
            for q in range(4):
                err = max(err, abs(rows[lane][q] - ref[2 * t + (q & 1), r0 + oi, x0 + g + 8 * (q >> 1)]))
    print("conv0 (pair-duplicated frame, one MMA per filter row): max err", err)
    assert err < 1e-9


def check_conv1(p):
    x = rng.random((32, 32, 8))
    E0 = plane(x)
    ref = conv_ref(x.transpose(2, 0, 1), p["w1"])
    w = [[wfrag(p, F_C1 + ky * 2 + h) for h in range(2)] for ky in range(3)]
    x0, r0 = 16, 8
    def loadA(i):
        aA = [(r0 + i, x0 + lr + 8 * (lj & 1) + (lj >> 1)) for _, lj, lr, _, _ in lanes()]
        aB = [(r0 + i, x0 + lr + 8 * (lj & 1) + 2) for _, lj, lr, _, _ in lanes()]
        f0 = ldsm(lambda a: E0[a[0], a[1]], aA, 4, False)
        f1 = ldsm(lambda a: E0[a[0], a[1]], aB, 2, False)
        f1 = [[r[0], r[1], ZERO, ZERO] for r in f1]
        return [f0, f1]
    acc = slide(4, 2, w, loadA)
    err = 0.0
    for oi, rows in acc.items():
        for lane, _, _, g, t in lanes():
            for q in range(4):
                err = max(err, abs(rows[lane][q] - ref[2 * t + (q & 1), r0 + oi, x0 + g + 8 * (q >> 1)]))
    print("conv1 (tap-paired, sliding): max err", err)
    assert err < 1e-9


def check_dec0(p):
    """cat(e0, up(o1)) 16 -> 8 on 32x32, one k16 step per tap, upsample as address map"""
    e0, o1 = rng.random((32, 32, 8)), rng.random((16, 16, 8))
    E0, O1 = plane(e0), plane(o1)
    cat = np.concatenate((e0, np.repeat(np.repeat(o1, 2, 0), 2, 1)), axis=2)
    ref = conv_ref(cat.transpose(2, 0, 1), p["d0"])
    w = [[wfrag(p, F_D0 + ky * 3 + kx) for kx in range(3)] for ky in range(3)]
    x0, r0 = 16, 4
    def loadA(i):
        out = []
        sy = (r0 + i + 1) >> 1
        for kx in range(3):
            addrs = []
            for _, lj, lr, _, _ in lanes():
                pixoff, chunk = lr + 8 * (lj & 1), lj >> 1
                addrs.append(("E", r0 + i, x0 + pixoff + kx) if chunk == 0 else ("O", sy, (x0 + pixoff + kx + 1) >> 1))
            out.append(ldsm(lambda a: (E0 if a[0] == "E" else O1)[a[1], a[2]], addrs, 4, False))
        return out
    acc = slide(4, 3, w, loadA)
    err = 0.0
    for oi, rows in acc.items():
        for lane, _, _, g, t in lanes():
            for q in range(4):
                err = max(err, abs(rows[lane][q] - ref[2 * t + (q & 1), r0 + oi, x0 + g + 8 * (q >> 1)]))
    print("dec0 (concat + nearest upsample by address map): max err", err)
    assert err < 1e-9


def wgrad_slide(R, y0, loadA, loadB):
    acc = [[np.zeros(4) for _ in range(32)] for _ in range(3)]
    b = {}
    for i in range(R + 2):
        a = loadA(y0 + i)
        if i < R:
            b[i % 3] = loadB(y0 + i)
        for ky in range(3):
            y = i - ky
            if 0 <= y < R:
                mma(acc[ky], a, b[y % 3])
    return acc


def check_dec0_wgrad(p):
    """dW[co][ci][ky][kx] = sum in[y+ky-1][x+kx-1][ci] * dY[y][x][co]; triples (src, kx group) over both strips and row halves"""
    e0, o1, dy = rng.random((32, 32, 8)), rng.random((16, 16, 8)), rng.standard_normal((32, 32, 8))
    E0, O1, DO0 = plane(e0), plane(o1), plane(dy)
    cat = np.concatenate((e0, np.repeat(np.repeat(o1, 2, 0), 2, 1)), axis=2)
    catp = np.zeros((34, 34, 16)); catp[1:-1, 1:-1] = cat
    ref = np.zeros((8, 16, 3, 3))
    for ky in range(3):
        for kx in range(3):
            ref[:, :, ky, kx] = np.einsum("yxo,yxc->oc", dy, catp[ky:ky + 32, kx:kx + 32])
    refb = dy.sum((0, 1))
    got = np.zeros((8, 16, 3, 3)); gotb = np.zeros(8)
    ONES = (1.0, 1.0)
    for tr in range(4):
        src, kxg = tr >> 1, tr & 1
        for kh in range(2):
            for s in range(2):
                x0 = 16 * s
                def loadB(y):
                    addrs = [(y + 1, 1 + x0 + lr + 8 * (lj & 1)) for _, lj, lr, _, _ in lanes()]
                    return ldsm(lambda a: DO0[a[0], a[1]], addrs, 2, True)
                def loadA(i):
                    if kxg == 0:
                        addrs = []
                        for _, lj, lr, _, _ in lanes():
                            tsel, tpix = lj & 1, lr + 8 * (lj >> 1)
                            v = x0 + tpix + tsel
                            addrs.append(("E", i, v) if src == 0 else ("O", (i + 1) >> 1, (v + 1) >> 1))
                        return ldsm(lambda a: (E0 if a[0] == "E" else O1)[a[1], a[2]], addrs, 4, True)
                    addrs = []
                    for _, lj, lr, _, _ in lanes():
                        v = x0 + lr + 8 * (lj & 1) + 2
                        addrs.append(("E", i, v) if src == 0 else ("O", (i + 1) >> 1, (v + 1) >> 1))
                    r = ldsm(lambda a: (E0 if a[0] == "E" else O1)[a[1], a[2]], addrs, 2, True)
                    out = []
                    for lane, _, _, g, _ in lanes():
                        one = ONES if (g == 0 and src == 0) else ZERO
                        out.append([r[lane][0], one, r[lane][1], one])
                    return out
                acc = wgrad_slide(16, 16 * kh, loadA, loadB)
                for ky in range(3):
                    for lane, _, _, g, t in lanes():
                        for q in range(4):
                            m, co = g + 8 * (q >> 1), 2 * t + (q & 1)
                            v = acc[ky][lane][q]
                            if kxg == 0:
                                got[co, src * 8 + (m & 7), ky, m >> 3] += v
                            elif m < 8:
                                got[co, src * 8 + m, ky, 2] += v
                            elif m == 8 and src == 0 and ky == 0:
                                gotb[co] += v
    print("dec0 wgrad (ldmatrix.trans operands, sliding triples): max err", np.abs(got - ref).max(), np.abs(gotb - refb).max())
    assert np.abs(got - ref).max() < 1e-8 and np.abs(gotb - refb).max() < 1e-8


def check_m0_wgrad_rgb(p):
    """masker.0 weight gradient, RGB triple (pair-duplicated frame rows), one band"""
    band = 1
    x = rng.random((64, 64, 3))
    dp = rng.standard_normal((18, 64, 16))                       # gradient band rows r = mask rows 16*band - 1 + r
    XB = np.zeros((20, 66, 8))
    full = pairdup(x)                                             # rows = haloed frame rows
    for rho in range(20):
        hv = 16 * band - 1 + rho
        if 0 <= hv < 66:
            XB[rho] = full[hv]
    DP = np.zeros((2, 18, 66, 8))
    for pl in range(2):
        DP[pl, :, 1:65] = dp[:, :, 8 * pl:8 * pl + 8]
    xp = np.zeros((66, 66, 3)); xp[1:65, 1:65] = x
    ref = np.zeros((16, 3, 3, 3))
    for r in range(1, 17):
        ya = 16 * band - 1 + r
        for ky in range(3):
            for kx in range(3):
                ref[:, :, ky, kx] += np.einsum("xo,xc->oc", dp[r], xp[ya + ky, kx:kx + 64])
    got = np.zeros((16, 3, 3, 3))
    for nt in range(2):
        for kh in range(2):
            r0 = 1 + 8 * kh
            for s in range(4):
                x0 = 16 * s
                def loadB(r):
                    addrs = [(r, 1 + x0 + lr + 8 * (lj & 1)) for _, lj, lr, _, _ in lanes()]
                    return ldsm(lambda a: DP[nt, a[0], a[1]], addrs, 2, True)
                def loadA(i):
                    addrs = [(i, x0 + (lr + 8 * (lj >> 1)) + 2 * (lj & 1)) for _, lj, lr, _, _ in lanes()]
                    return ldsm(lambda a: XB[a[0], a[1]], addrs, 4, True)
                acc = wgrad_slide(8, r0, loadA, loadB)
                for ky in range(3):
                    for lane, _, _, g, t in lanes():
                        for q in range(4):
                            m, co = g + 8 * (q >> 1), nt * 8 + 2 * t + (q & 1)
                            kx, c = m >> 2, m & 3
                            if kx < 3 and c < 3:
                                got[co, c, ky, kx] += acc[ky][lane][q]
    print("masker.0 wgrad, RGB triple: max err", np.abs(got - ref).max())
    assert np.abs(got - ref).max() < 1e-8


def check_m2(p):
    """masker.2 weight gradient and input gradient with the hand-gathered d-logit fragments, one band"""
    band = 2
    m0 = rng.standard_normal((18, 64, 16))                        # band rows r <-> mask rows 16*band - 1 + r (all inside)
    dl = rng.standard_normal((20, 64))                            # rows rho <-> mask rows 16*band - 2 + rho
    M0 = np.zeros((2, 18, 66, 8))
    for pl in range(2):
        M0[pl, :, 1:65] = m0[:, :, 8 * pl:8 * pl + 8]
    DL = np.zeros((20, DLP)); DL[:, 1:65] = dl
    # reference: dW2[ci][ky][kx] = sum over interior mask rows y (rho 2..17), x of m0[y+ky-1][x+kx-1][ci] * dl[y][x]
    m0p = np.zeros((18, 66, 16)); m0p[:, 1:65] = m0
    ref = np.zeros((16, 3, 3))
    for rho in range(2, 18):
        for ky in range(3):
            r = rho - 2 + ky                                      # band row of m0: mask row (16b-2+rho)+ky-1 = 16b-1 + r
            for kx in range(3):
                ref[:, ky, kx] += np.einsum("x,xc->c", dl[rho], m0p[r, kx:kx + 64])
    acc2 = [[np.zeros(4) for _ in range(32)] for _ in range(2)]
    for pg in range(72):
        r, s = pg >> 2, pg & 3
        addrs = [((lj & 1), r, 1 + 16 * s + lr + 8 * (lj >> 1)) for _, lj, lr, _, _ in lanes()]
        a = ldsm(lambda q: M0[q[0], q[1], q[2]], addrs, 4, True)
        b0, b1 = [], []
        for lane, _, _, g, t in lanes():
            ky, kx = g // 3, g % 3
            rho = r - ky + 2
            if 2 <= rho < 18:
                c = 16 * s + 2 * t - kx + 2
                b0.append([(DL[rho, c], DL[rho, c + 1]), (DL[rho, c + 8], DL[rho, c + 9])])
            else:
                b0.append([ZERO, ZERO])
            rho = r
            if g == 0 and 2 <= rho < 18:
                c = 16 * s + 2 * t
                b1.append([(DL[rho, c], DL[rho, c + 1]), (DL[rho, c + 8], DL[rho, c + 9])])
            else:
                b1.append([ZERO, ZERO])
        mma(acc2[0], a, b0)
        mma(acc2[1], a, b1)
    got = np.zeros((16, 9))
    for lane, _, _, g, t in lanes():
        for q in range(4):
            ci, col = g + 8 * (q >> 1), 2 * t + (q & 1)
            got[ci, col] += acc2[0][lane][q]
            if col == 0:
                got[ci, 8] += acc2[1][lane][q]
    print("masker.2 wgrad: max err", np.abs(got.reshape(16, 3, 3) - ref).max())
    assert np.abs(got.reshape(16, 3, 3) - ref).max() < 1e-8
    # input gradient: dm0[r][x][ci] = sum_{ky,kx} W2[ci][ky][kx] * dl[mask row - ky + 1][x - kx + 1]
    w2 = p["m2"].reshape(16, 3, 3)
    dlp = np.zeros((22, 66)); dlp[1:21, 1:65] = dl               # dlp[rho + 1][x + 1]
    refd = np.zeros((18, 64, 16))
    for r in range(18):
        rho_c = r + 1                                             # mask row 16b-1+r = 16b-2 + (r+1)
        for ky in range(3):
            for kx in range(3):
                refd[r] += np.einsum("x,c->xc", dlp[rho_c - ky + 1 + 1, (1 - kx + 1):(1 - kx + 1) + 64], w2[:, ky, kx])
    wd = [wfrag(p, B_M2D + nt) for nt in range(2)]
    err = 0.0
    for pg in range(72):
        r, s = pg >> 2, pg & 3
        a = []
        for lane, _, _, g, t in lanes():
            ta, tb = 2 * t, 2 * t + 1
            q0 = lambda off: DL[r + ta // 3, 16 * s + g + ta % 3 + off]
            q1 = lambda off: DL[r + tb // 3, 16 * s + g + tb % 3 + off]
            q8 = lambda off: DL[r + 2, 16 * s + g + 2 + off]
            a.append([(q0(0), q1(0)), (q0(8), q1(8)), ((q8(0), 0.0) if t == 0 else ZERO), ((q8(8), 0.0) if t == 0 else ZERO)])
        for nt in range(2):
            c = [np.zeros(4) for _ in range(32)]
            mma(c, a, wd[nt])
            for lane, _, _, g, t in lanes():
                for q in range(4):
                    err = max(err, abs(c[lane][q] - refd[r, 16 * s + g + 8 * (q >> 1), nt * 8 + 2 * t + (q & 1)]))
    print("masker.2 dgrad: max err", err)
    assert err < 1e-8


def check_m0_dgrad(p):
    """masker.0 input gradient into the upsampled o0 channels (16 co -> 8 ci), one k16 step per tap', rows of the band"""
    dp = rng.standard_normal((18, 64, 16))
    DP = np.zeros((2, 18, 66, 8))
    for pl in range(2):
        DP[pl, :, 1:65] = dp[:, :, 8 * pl:8 * pl + 8]
    w0 = p["m0"]
    dpp = np.zeros((18, 66, 16)); dpp[:, 1:65] = dp
    ref = np.zeros((16, 64, 8))                                   # output rows o <-> band rows o+1
    for o in range(16):
        for ky in range(3):
            for kx in range(3):
                # d in[y][x] += dout[y - ky + 1][x - kx + 1] * W[co][ci][ky][kx]
                rr = (o + 1) - ky + 1
                ref[o] += np.einsum("xo,oc->xc", dpp[rr, (1 - kx + 1):(1 - kx + 1) + 64], w0[:, 3:, ky, kx])
    w = [[wfrag(p, B_M0D + ky * 3 + kx) for kx in range(3)] for ky in range(3)]
    x0 = 32
    def loadA(i):
        out = []
        for kx in range(3):
            addrs = [((lj >> 1), i, x0 + lr + 8 * (lj & 1) + kx) for _, lj, lr, _, _ in lanes()]
            out.append(ldsm(lambda a: DP[a[0], a[1], a[2]], addrs, 4, False))
        return out
    acc = slide(16, 3, w, loadA)
    err = 0.0
    for oi, rows in acc.items():
        for lane, _, _, g, t in lanes():
            for q in range(4):
                err = max(err, abs(rows[lane][q] - ref[oi, x0 + g + 8 * (q >> 1), 2 * t + (q & 1)]))
    print("masker.0 dgrad into up(o0): max err", err)
    assert err < 1e-8


def dgrad_ref(dy_chw, w):
    """d input of a 3x3 s1 p1 conv: dx[ci][y][x] = sum dy[co][y-ky+1][x-kx+1] * w[co][ci][ky][kx]"""
    C, H, W = dy_chw.shape
    dp = np.zeros((C, H + 2, W + 2)); dp[:, 1:-1, 1:-1] = dy_chw
    out = np.zeros((w.shape[1], H, W))
    for ky in range(3):
        for kx in range(3):
            out += np.einsum("oc,ohw->chw", w[:, :, ky, kx], dp[:, 2 - ky:2 - ky + H, 2 - kx:2 - kx + W])
    return out


def check_critic_dgrads(p):
    """features.3 (8 -> 8) and features.0 (8 -> 3) input gradients: tap-paired sliding conv of the haloed gradient plane with
    the rotated filters (pack steps B_C1D, B_C0D); features.10 (16 -> 8): one k16 step per tap' (B_C3D)"""
    for name, base, wkey, cin in (("features.3", B_C1D, "w1", 8), ("features.0", B_C0D, "w0", 3)):
        dy = rng.standard_normal((32, 32, 8))
        DY = plane(dy)
        ref = dgrad_ref(dy.transpose(2, 0, 1), p[wkey])
        w = [[wfrag(p, base + ky * 2 + h) for h in range(2)] for ky in range(3)]
        x0, r0 = 16, 8
        def loadA(i):
            aA = [(r0 + i, x0 + lr + 8 * (lj & 1) + (lj >> 1)) for _, lj, lr, _, _ in lanes()]
            aB = [(r0 + i, x0 + lr + 8 * (lj & 1) + 2) for _, lj, lr, _, _ in lanes()]
            f0 = ldsm(lambda a: DY[a[0], a[1]], aA, 4, False)
            f1 = ldsm(lambda a: DY[a[0], a[1]], aB, 2, False)
            return [f0, [[r[0], r[1], ZERO, ZERO] for r in f1]]
        acc = slide(4, 2, w, loadA)
        err = 0.0
        for oi, rows in acc.items():
            for lane, _, _, g, t in lanes():
                for q in range(4):
                    c = 2 * t + (q & 1)
                    want = ref[c, r0 + oi, x0 + g + 8 * (q >> 1)] if c < cin else 0.0
                    err = max(err, abs(rows[lane][q] - want))
        print(f"{name} dgrad (rotated filter, tap-paired): max err", err)
        assert err < 1e-8
    dy = rng.standard_normal((8, 8, 16))
    DY = np.stack([plane(dy[:, :, :8]), plane(dy[:, :, 8:])])
    ref = dgrad_ref(dy.transpose(2, 0, 1), p["w3"])
    err = 0.0
    for mt in range(4):
        acc = [np.zeros(4) for _ in range(32)]
        for tp in range(9):
            ky, kx = tp // 3, tp % 3
            addrs = [((lj >> 1), 2 * mt + (lj & 1) + ky, lr + kx) for _, lj, lr, _, _ in lanes()]
            mma(acc, ldsm(lambda a: DY[a[0], a[1], a[2]], addrs, 4, False), wfrag(p, B_C3D + tp))
        for lane, _, _, g, t in lanes():
            for q in range(4):
                err = max(err, abs(acc[lane][q] - ref[2 * t + (q & 1), 2 * mt + (q >> 1), g]))
    print("features.10 dgrad (16 -> 8): max err", err)
    assert err < 1e-8


if __name__ == "__main__":
    p = dict(w0=rng.standard_normal((8, 3, 3, 3)), w1=rng.standard_normal((8, 8, 3, 3)), w2=rng.standard_normal((8, 8, 3, 3)),
             w3=rng.standard_normal((16, 8, 3, 3)), d0=rng.standard_normal((8, 16, 3, 3)), d1=rng.standard_normal((8, 16, 3, 3)),
             d2=rng.standard_normal((8, 24, 3, 3)), d3=rng.standard_normal((16, 48, 3, 3)), m0=rng.standard_normal((16, 11, 3, 3)),
             m2=rng.standard_normal((1, 16, 3, 3)))
    check_conv0(p)
    check_conv1(p)
    check_dec0(p)
    check_dec0_wgrad(p)
    check_m0_wgrad_rgb(p)
    check_m2(p)
    check_m0_dgrad(p)
    check_critic_dgrads(p)
    print("all emulated phases agree with the numpy references")
