"""The value-range pruning in k_fold relies on |simplex3| <= 1 (so |fBm * noise_scale| <= amplitude).
Checked here two ways: an upper envelope over ALL gradient assignments (independent numpy
restatement of the falloff / skew geometry) and the oracle's actual noise on dense samples."""
import numpy as np


def _envelope(P):
    F3, G3 = 1.0 / 3.0, 1.0 / 6.0
    x, y, z = P[:, 0], P[:, 1], P[:, 2]
    f = F3 * (x + y + z)
    x0, y0, z0 = np.floor(x + f), np.floor(y + f), np.floor(z + f)
    g = G3 * (x0 + y0 + z0)
    x0, y0, z0 = x - (x0 - g), y - (y0 - g), z - (z0 - g)
    xy, yz, xz = x0 >= y0, y0 >= z0, x0 >= z0
    i1, j1, k1 = xy & xz, (~xy) & yz, (~xz) & (~yz)
    i2, j2, k2 = xy | xz, (~xy) | yz, ~(xz & yz)
    ds = [(x0, y0, z0), (x0 - i1 + G3, y0 - j1 + G3, z0 - k1 + G3), (x0 - i2 + F3, y0 - j2 + F3, z0 - k2 + F3),
          (x0 - 0.5, y0 - 0.5, z0 - 0.5)]
    tot = 0.0
    for a, b, c in ds:
        t = np.maximum(0.6 - a * a - b * b - c * c, 0.0)
        comps = np.sort(np.abs(np.stack([a, b, c], 1)), axis=1)
        tot = tot + t ** 4 * (comps[:, 1] + comps[:, 2])  # max over the 12 edge gradients of |g . d|
    return 32.69428253173828125 * tot


def test_simplex3_envelope_over_all_gradients_is_at_most_one():
    rng = np.random.default_rng(0)
    m = 0.0
    for _ in range(6):
        m = max(m, _envelope(rng.uniform(0.0, 3.0, (1_000_000, 3))).max())
    g = np.linspace(0.0, 1.5, 120)
    m = max(m, _envelope(np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)).max())
    assert 0.99 < m <= 1.0 + 1e-6, m


def test_oracle_fbm_stays_within_its_theoretical_amplitude(oracle):
    rng = np.random.default_rng(1)
    pts = rng.uniform(-50.0, 50.0, (20000, 3)).astype(np.float32)
    s = np.array([oracle.simplex3(float(p[0]), float(p[1]), float(p[2]), 3) for p in pts[:8000]])
    assert np.abs(s).max() <= 1.0 + 1e-5
    gain, octaves = 0.546, 5
    inherent = (1.0 - gain ** octaves) / (1.0 - gain)
    f = np.array([oracle.fbm3(float(p[0]), float(p[1]), float(p[2]), 2.0, gain, octaves, 0) for p in pts[:8000]])
    assert np.abs(f).max() <= inherent * (1.0 + 1e-5)
