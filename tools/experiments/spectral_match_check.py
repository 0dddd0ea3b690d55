"""Design check for DESIGN.md section 8.1 (not product code, CPU only): the 120 shift / reversal variants of processSC.m:24-31
from per-ring spectra.  For a query image x and a DB image h (60 sectors x 20 rings, L2-normalised),
    corr_x[s] = sum_{c,r} x[(c+s)%60][r] h[c][r] = irfft_f( sum_r X_r[f] conj(H_r[f]) )[s]
    corr_y[s] (y = sector reversal of x, Y_r[f] = conj(X_r[f])) = irfft_f( sum_r conj(X_r[f]) conj(H_r[f]) )[s]
so a pair needs 31 complex dot products of length 20 per base (4 real bilinear forms per frequency) and two inverse
transforms 31 -> 60, instead of 120 dot products of length 1200.  The script checks the identity in fp64 and emulates the
arithmetic a tensor-core version would use (operands split into fp16 hi + lo, three products, fp32 accumulation; the
cross-spectra split again for the inverse transform) to see what precision survives."""
import numpy as np

rng = np.random.default_rng(20261017)


def make_images(n):
    img = rng.gamma(2.0, 1.5, size=(n, 60, 20)) * (rng.random((n, 60, 20)) < 0.35)   # ~35 % occupied bins
    spikes = rng.random(n) < 0.1                                                      # some spiky signatures
    img[spikes] *= (rng.random((spikes.sum(), 60, 20)) < 0.02) * 20 + 1e-3
    nrm = np.sqrt((img ** 2).sum(axis=(1, 2), keepdims=True))
    return img / nrm


def direct(x, h):
    best = -np.inf
    for base in (x, np.roll(x[::-1], 1, axis=0)):            # y[c] = x[(60 - c) % 60]
        for s in range(60):
            best = max(best, float((np.roll(base, -s, axis=0) * h).sum()))
    return (1.0 - best) / 2.0


def split16(a):
    hi = a.astype(np.float16)
    lo = (a - hi.astype(np.float64)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def dot3(a, b):
    """sum over the last axis of a*b with the 3-term fp16 split and fp32 accumulation (hi*lo + lo*hi + hi*hi)"""
    ah, al = split16(a)
    bh, bl = split16(b)
    return ((ah * bl).sum(-1, dtype=np.float32) + (al * bh).sum(-1, dtype=np.float32) + (ah * bh).sum(-1, dtype=np.float32)).astype(np.float64)


def spectral(x, h, emulate):
    X = np.fft.rfft(x, axis=0)          # (31, 20)
    H = np.fft.rfft(h, axis=0)
    if not emulate:
        Sx = (X * np.conj(H)).sum(1)
        Sy = (np.conj(X) * np.conj(H)).sum(1)
        cx, cy = np.fft.irfft(Sx, 60), np.fft.irfft(Sy, 60)
    else:
        sc = 64.0
        xr, xi, hr, hi_ = X.real * sc, X.imag * sc, H.real * sc, H.imag * sc
        p1, p2, p3, p4 = dot3(xr, hr), dot3(xi, hi_), dot3(xi, hr), dot3(xr, hi_)    # 4 real forms per frequency
        Sx = (p1 + p2) + 1j * (p3 - p4)
        Sy = (p1 - p2) - 1j * (p3 + p4)
        # inverse transform as a real GEMM: corr[s] = sum_k S_k W[k, s], operands split again
        f = np.arange(31)[:, None]
        s = np.arange(60)[None, :]
        w = np.where((f == 0) | (f == 30), 1.0, 2.0) / 60.0
        Wc, Ws = w * np.cos(2 * np.pi * f * s / 60), -w * np.sin(2 * np.pi * f * s / 60)
        out = []
        for S in (Sx, Sy):
            a = np.concatenate([S.real, S.imag])                       # (62,)
            Wm = np.concatenate([Wc, Ws], axis=0)                      # (62, 60)
            out.append(dot3(np.broadcast_to(a, (60, 62)), Wm.T) / (sc * sc))
        cx, cy = out
    return (1.0 - max(cx.max(), cy.max())) / 2.0


def main():
    q, d = make_images(48), make_images(48)
    e_exact = e_emul = 0.0
    for i in range(48):
        for j in range(0, 48, 3):
            ref = direct(q[i], d[j])
            e_exact = max(e_exact, abs(spectral(q[i], d[j], False) - ref))
            e_emul = max(e_emul, abs(spectral(q[i], d[j], True) - ref))
    print(f"pairs checked: {48 * 16}")
    print(f"max |d_spectral - d_direct|, fp64:                           {e_exact:.2e}")
    print(f"max |d_spectral - d_direct|, fp16 hi/lo x3 + fp32 accumulate: {e_emul:.2e}   (bar 1e-5)")
    mac_direct = 120 * 1200
    mac_halved = 2 * 30 * 600 * 2
    mac_spec = 31 * 20 * 4 + 2 * 62 * 60
    print(f"multiply-adds per pair and channel: direct {mac_direct}, even/odd halving (today) {mac_halved}, spectral {mac_spec} "
          f"({mac_halved / mac_spec:.1f}x fewer than today; x3 each for the split)")


if __name__ == "__main__":
    main()
