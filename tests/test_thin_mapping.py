"""The fragment bookkeeping of the warp-MMA thin-block kernel (vidsitu_b200/csrc/bottleneck_thin_sm100.cu), replayed on
the CPU: a numpy model of the mma.sync m16n8k16 / m16n8k8 register fragments is driven with exactly the index
expressions the kernel uses (the channel permutation of conv a's K order and conv c's N order, the 8-channel K slices
of conv b paired into K = 16 steps, the b accumulator re-used as conv c's A fragment, the residual taken from the x
fragments) and must reproduce a -> b -> c + residual computed directly.  No GPU, no product code: this pins the
mapping the kernel's comments describe."""
import numpy as np
import pytest


def mma(acc, a, b, k16=True):
    """D += A . B on per-lane fragments: a[lane] = (a0, a1[, a2, a3]) pairs, b[lane] = (b0[, b1]) pairs, acc[lane][4]."""
    K = 16 if k16 else 8
    A = np.zeros((16, K))
    B = np.zeros((K, 8))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        A[g, 2 * t:2 * t + 2] = a[lane][0]
        A[g + 8, 2 * t:2 * t + 2] = a[lane][1]
        B[2 * t:2 * t + 2, g] = b[lane][0]
        if k16:
            A[g, 2 * t + 8:2 * t + 10] = a[lane][2]
            A[g + 8, 2 * t + 8:2 * t + 10] = a[lane][3]
            B[2 * t + 8:2 * t + 10, g] = b[lane][1]
    Dm = A @ B
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        acc[lane][0] += Dm[g, 2 * t]
        acc[lane][1] += Dm[g, 2 * t + 1]
        acc[lane][2] += Dm[g + 8, 2 * t]
        acc[lane][3] += Dm[g + 8, 2 * t + 1]


@pytest.mark.parametrize("D,KT", [(8, 1), (8, 3), (16, 1), (16, 3)])
def test_identity_block_fragment_mapping(D, KT):
    rng = np.random.default_rng(D * 10 + KT)
    C, NQ, NTD = 4 * D, D // 8, D // 8
    NSL = 9 * NQ
    NP = (NSL + 1) // 2
    W, R = 6, 3
    a_px = (R + 2) * W
    x = rng.standard_normal((KT, a_px, C))
    wa = rng.standard_normal((D, KT, C))
    wb = rng.standard_normal((D, 9, D))
    wc = rng.standard_normal((C, D))
    a_ref = np.einsum("tpc,dtc->pd", x, wa)
    # conv a: K step (tap, q, h) of thread t covers channels 8 (4 q + t) + 4 h + {0, 1} (a0 / a1) and + {2, 3} (a2 / a3)
    abuf = np.zeros((R + 2, W + 2, D))
    for tile in range((a_px + 15) // 16):
        acc = [[[0.0] * 4 for _ in range(32)] for _ in range(NTD)]
        for tap in range(KT):
            for q in range(NQ):
                for h in range(2):
                    for nt in range(NTD):
                        a, b = [], []
                        for lane in range(32):
                            g, t = lane >> 2, lane & 3
                            pg, ph = min(tile * 16 + g, a_px - 1), min(tile * 16 + g + 8, a_px - 1)
                            ch0 = 8 * (4 * q + t)
                            vg = [x[tap, pg, ch0 + 2 * i:ch0 + 2 * i + 2] for i in range(4)]
                            vh = [x[tap, ph, ch0 + 2 * i:ch0 + 2 * i + 2] for i in range(4)]
                            a.append([vg[2 * h], vh[2 * h], vg[2 * h + 1], vh[2 * h + 1]])
                            base = ch0 + 4 * h
                            b.append([wa[8 * nt + g, tap, base:base + 2], wa[8 * nt + g, tap, base + 2:base + 4]])
                        mma(acc[nt], a, b)
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            for half in range(2):
                px = tile * 16 + g + 8 * half
                if px < a_px:
                    rr, col = divmod(px, W)
                    for nt in range(NTD):
                        abuf[rr, col + 1, 8 * nt + 2 * t] = acc[nt][lane][2 * half]
                        abuf[rr, col + 1, 8 * nt + 2 * t + 1] = acc[nt][lane][2 * half + 1]
    assert np.allclose(abuf[:, 1:-1, :].reshape(a_px, D), a_ref)

    bc_px = R * W
    b_ref = np.zeros((R, W, D))
    for rr in range(R):
        for col in range(W):
            for tap in range(9):
                b_ref[rr, col] += wb[:, tap, :] @ abuf[rr + tap // 3, col + tap % 3]
    xres = rng.standard_normal(((R + 2) * W, C))
    out_ref = np.einsum("rwd,cd->rwc", b_ref, wc) + xres.reshape(R + 2, W, C)[1:R + 1]
    out = np.zeros((R, W, C))
    for tile in range((bc_px + 15) // 16):
        def centre(lane, half):
            return divmod(min(tile * 16 + (lane >> 2) + 8 * half, bc_px - 1), W)
        # conv b: K slices s = tap * NQ + q of 8 channels, two per MMA
        accb = [[[0.0] * 4 for _ in range(32)] for _ in range(NTD)]
        for pr in range(NP):
            s0, s1 = 2 * pr, 2 * pr + 1
            for nt in range(NTD):
                a, b = [], []
                for lane in range(32):
                    g, t = lane >> 2, lane & 3

                    def afrag(s, half):
                        tap, q = s // NQ, s % NQ
                        rr, col = centre(lane, half)
                        return abuf[rr + tap // 3, col + tap % 3, 8 * q + 2 * t:8 * q + 2 * t + 2]

                    def bfrag(s):
                        tap, q = s // NQ, s % NQ
                        return wb[8 * nt + g, tap, 8 * q + 2 * t:8 * q + 2 * t + 2]
                    z = np.zeros(2)
                    a.append([afrag(s0, 0), afrag(s0, 1), afrag(s1, 0) if s1 < NSL else z, afrag(s1, 1) if s1 < NSL else z])
                    b.append([bfrag(s0), bfrag(s1) if s1 < NSL else z])
                mma(accb[nt], a, b)
        # the accumulator fragment of b IS the A fragment of c
        pb = [[[np.array(accb[nt][lane][0:2]), np.array(accb[nt][lane][2:4])] for lane in range(32)] for nt in range(NTD)]
        for q in range(NQ):
            for i in range(4):
                accc = [[0.0] * 4 for _ in range(32)]
                a, b = [], []
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    co = 8 * (4 * q + (g >> 1)) + 2 * i + (g & 1)   # column n = g of tile (q, i)
                    if D >= 16:
                        a.append([pb[0][lane][0], pb[0][lane][1], pb[1][lane][0], pb[1][lane][1]])
                        b.append([wc[co, 2 * t:2 * t + 2], wc[co, 8 + 2 * t:8 + 2 * t + 2]])
                    else:
                        a.append([pb[0][lane][0], pb[0][lane][1]])
                        b.append([wc[co, 2 * t:2 * t + 2]])
                mma(accc, a, b, k16=D >= 16)
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    for half in range(2):
                        px = tile * 16 + g + 8 * half
                        if px < bc_px:
                            rr, col = divmod(px, W)
                            ch = 8 * (4 * q + t) + 2 * i          # = register i of the thread's x piece q: the residual
                            res = xres[(rr + 1) * W + col, ch:ch + 2]
                            out[rr, col, ch] = accc[lane][2 * half] + res[0]
                            out[rr, col, ch + 1] = accc[lane][2 * half + 1] + res[1]
    assert np.allclose(out, out_ref)
