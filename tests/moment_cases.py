"""Adversarial addend sequences for the serial-order moment sums (KMomentsSerial): the kernel must return the value of the
dependent chain  acc = 0; for a in addends: acc += a  bit for bit, whatever the data does."""
import numpy as np


def serial_sum(a):
    acc = np.float64(0.0)
    out = np.add.accumulate(np.concatenate([[acc], np.asarray(a, np.float64)]))      # accumulate is the sequential chain
    return out[-1]


def cases(seed=0):
    rng = np.random.default_rng(seed)
    n = 5000
    out = {}
    out["uniform"] = rng.random(n)
    out["normal_cancelling"] = rng.standard_normal(n)
    out["log_uniform"] = np.exp(rng.uniform(-40, 40, n)) * rng.choice([-1.0, 1.0], n)
    out["growing"] = np.exp(np.linspace(-30, 30, n))
    out["shrinking_alternating"] = np.exp(np.linspace(30, -30, n)) * np.where(np.arange(n) % 2 == 0, 1.0, -1.0)
    base = np.full(n, 1.0); base[0] = 2.0 ** 30
    out["big_then_ones"] = base
    # ties: a running sum of 2^53-scale integers plus exact half-ulp addends (round-to-even decides every step)
    t = np.full(n, 1.0); t[0] = 2.0 ** 53; t[1::2] = 1.0; t[2::2] = 3.0
    out["ties_integer"] = t
    h = np.full(n, 2.0 ** -53); h[0] = 1.0
    out["ties_half_ulp_of_one"] = h
    h2 = np.where(rng.random(n) < 0.5, 2.0 ** -53, -(2.0 ** -54)); h2[0] = 1.0 + 2.0 ** -52
    out["ties_mixed_signs"] = h2
    m = rng.integers(0, 4, n).astype(np.float64) * 2.0 ** -53 * rng.choice([-1.0, 1.0], n); m[0] = 1.5
    out["ties_multiples_of_half_ulp"] = m
    z = np.zeros(n); z[n // 2] = 1e-300; z[n // 2 + 1] = -1e-300
    out["zeros_and_tiny"] = z
    out["all_zero"] = np.zeros(n)
    nz = np.zeros(n); nz[:] = -0.0
    out["all_negative_zero"] = nz
    sub = rng.integers(1, 1000, n).astype(np.float64) * 5e-324 * rng.choice([-1.0, 1.0], n)
    out["subnormals"] = sub
    big = np.full(n, 1e307); big[n // 2:] = -1e307
    out["near_overflow"] = big
    inf = rng.standard_normal(n); inf[n // 3] = np.inf
    out["inf"] = inf
    nan = rng.standard_normal(n); nan[n // 3] = np.inf; nan[2 * n // 3] = -np.inf
    out["inf_minus_inf"] = nan
    c = rng.standard_normal(n) * 1e-3; c[::100] = 1e3; c[50::100] = -1e3
    out["spikes_cancel"] = c
    out["binade_walk"] = np.where(np.arange(n) % 7 == 0, -0.9, 0.15) * (1.0 + rng.random(n) * 1e-12)
    out["long_random"] = rng.standard_normal(200000) + 0.01
    out["long_positive"] = rng.random(300000) * np.exp(rng.uniform(-20, 0, 300000))
    # long sequences for the tiled scan (tiles of 8192 addends are summarised in parallel under a GUESSED binade and applied only when the exact sum agrees)
    N = 150000
    tt = np.where(rng.random(N) < 0.5, 2.0 ** -53, 2.0 ** -52) * rng.choice([1.0, 1.0, -1.0], N); tt[0] = 1.0
    out["tiles_ties"] = tt                                               # half-ulp ties on a sum near 1: the parity maps carry across tiles
    out["tiles_binade_growth"] = np.exp(np.linspace(-25, 3, N)) * rng.random(N)      # crosses ~40 binades, some inside tiles, some at their edges
    edge = np.full(N, 2.0 ** -60); edge[0] = 1.0 - 2.0 ** -40; edge[8192 * 3 - 1] = 2.0 ** -40; edge[8192 * 7] = 2.0 ** -41
    out["tiles_cross_at_edges"] = edge                                   # the sum reaches exactly 1.0 on the last addend of a tile
    sw = rng.standard_normal(N) * 1e-6; sw[0] = 1.0; sw[N // 3] = -2.0; sw[2 * N // 3] = 2.5
    out["tiles_sign_changes"] = sw
    bo = np.where(np.arange(N) % 2 == 0, 0.75, -0.75) * (1.0 + rng.random(N) * 1e-9); bo[0] = 1.5
    out["tiles_big_offsets"] = bo                                        # addends as large as the sum: offsets of 2^52 ulps, the guards against overflowing them
    gi = rng.random(N) * 1e-3; gi[N // 2] = np.inf
    out["tiles_inf_inside"] = gi
    go = np.full(N, 2.0 ** -70); go[0] = 1.0 - 2.0 ** -53               # the exact sum stays below 1 for a long time while the unordered tile sums round to 1.0
    out["tiles_guess_off"] = go
    out["short_3"] = np.array([1.0, 2.0 ** -53, 2.0 ** -53])
    out["empty"] = np.zeros(0)
    out["exact_tile"] = rng.random(1024)
    out["tile_plus_one"] = rng.random(1025)
    return out
