"""The algebra `sc_match_tc_kernel` rests on (so_dso_place_recognition_b200/csrc/sc_match_tc.cu, header), checked on
the CPU against the plain definition of the 120 variants (processSC.m:24-31):

    e[u] = b[u] + b[u+30],  o[u] = b[u] - b[u+30]   (query base b, sector index mod 60)
    he[c] = h[c] + h[c+30], ho[c] = h[c] - h[c+30]  (DB row, c < 30)
    E[s] = sum_{c<30} e[c+s].he[c] = corr[s] + corr[s+30],   O[s] = sum_{c<30} o[c+s].ho[c] = corr[s] - corr[s+30]
    max_s corr[s] = max_{s<30} (E[s] + |O[s]|) / 2

for the forward base x and the sector-reversed base y[c] = x[(60-c) % 60], whose e / o are the reversals of x's."""
import numpy as np


def _corr_all(b, h):
    """corr[s] = sum_c b[(c+s) % 60] . h[c], b and h as (60, 20) sector-major images."""
    return np.array([(np.roll(b, -s, axis=0) * h).sum() for s in range(60)])


def test_even_odd_halving_gives_all_120_variants():
    rng = np.random.default_rng(3)
    for trial in range(20):
        x = rng.normal(size=(60, 20)) * (rng.random((60, 20)) < 0.4)
        h = rng.normal(size=(60, 20)) * (rng.random((60, 20)) < 0.4)
        y = x[(60 - np.arange(60)) % 60]
        he, ho = h[:30] + h[30:], h[:30] - h[30:]
        for b in (x, y):
            corr = _corr_all(b, h)
            e = b + np.roll(b, -30, axis=0)
            o = b - np.roll(b, -30, axis=0)
            E = np.array([(np.roll(e, -s, axis=0)[:30] * he).sum() for s in range(30)])
            O = np.array([(np.roll(o, -s, axis=0)[:30] * ho).sum() for s in range(30)])
            np.testing.assert_allclose(E, corr[:30] + corr[30:], atol=1e-10)
            np.testing.assert_allclose(O, corr[:30] - corr[30:], atol=1e-10)
            np.testing.assert_allclose((E + np.abs(O)).max() / 2, corr.max(), atol=1e-10)
        # e / o of the reversed base are the reversals of e / o of x (what the query operand preparation uses)
        ex, ox = x + np.roll(x, -30, axis=0), x - np.roll(x, -30, axis=0)
        ey, oy = y + np.roll(y, -30, axis=0), y - np.roll(y, -30, axis=0)
        rev = (60 - np.arange(60)) % 60
        np.testing.assert_array_equal(ey, ex[rev])
        np.testing.assert_array_equal(oy, ox[rev])


def test_binary_channel_stays_exact_in_e2m1():
    """0/1 inputs: e in {0, 1, 2}, o in {-1, 0, 1} (all representable in e2m1) and E, O are integers below 2^24."""
    rng = np.random.default_rng(4)
    x = (rng.random((60, 20)) < 0.5).astype(np.float64)
    e, o = x + np.roll(x, -30, axis=0), x - np.roll(x, -30, axis=0)
    assert set(np.unique(e)) <= {0.0, 1.0, 2.0} and set(np.unique(o)) <= {-1.0, 0.0, 1.0}
    assert 2 * 2 * 600 < 2 ** 24


def test_matches_oracle_numpy_matcher(oracle):
    """End to end on signatures: (1 - max(E + |O|)/2) / 2 equals the oracle's d over the 120 variants."""
    rng = np.random.default_rng(5)
    sig = rng.random((6, 2400)) * (rng.random((6, 2400)) < 0.3)
    sig[:, 1200:] = (sig[:, 1200:] > 0.15)
    dp, di = oracle.sc_match_numpy(sig, sig)
    for ch, ref in ((0, dp), (1, di)):
        a = sig[:, ch * 1200:(ch + 1) * 1200]
        a = a / np.linalg.norm(a, axis=1, keepdims=True)
        for i in range(6):
            x = a[i].reshape(60, 20)
            y = x[(60 - np.arange(60)) % 60]
            for j in range(6):
                h = a[j].reshape(60, 20)
                he, ho = h[:30] + h[30:], h[:30] - h[30:]
                best = -np.inf
                for b in (x, y):
                    e, o = b + np.roll(b, -30, axis=0), b - np.roll(b, -30, axis=0)
                    for s in range(30):
                        E = (np.roll(e, -s, axis=0)[:30] * he).sum()
                        O = (np.roll(o, -s, axis=0)[:30] * ho).sum()
                        best = max(best, (E + abs(O)) / 2)
                assert abs((1 - best) / 2 - ref[i, j]) < 1e-12
