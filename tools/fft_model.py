"""Design-time model of the register-resident Stockham FFT kernel (csrc/fft_pow2.cuh).

Mirrors the kernel's thread/register/shared-memory index math in NumPy so it can be checked on a
machine without a GPU:  (1) the pass recurrence and the fact that every pass reads the same register
positions t + m*N/R,  (2) the r2c post-process / c2r pre-process formulas,  (3) shared-memory bank
conflicts of the scattered store for a given padding rule.  Not part of the product.
"""
import itertools
import sys

import numpy as np


def radix_plan(N, R):
    """Radices (each <= R, powers of two) whose product is N; small radix last (cheap last-pass twiddles)."""
    plan = []
    rem = N
    while rem > 1:
        r = min(R, rem)
        plan.append(r)
        rem //= r
    # order: largest first, remainder last
    return plan


def stockham(x, radices, R, sign=-1):
    N = len(x)
    T = N // R
    v = np.array([[x[t + m * T] for m in range(R)] for t in range(T)], dtype=complex)
    Ns = 1
    for p, r in enumerate(radices):
        nb = R // r
        out = np.zeros(N, dtype=complex)
        last = p == len(radices) - 1
        for t in range(T):
            for b in range(nb):
                j = t + b * T
                a = j % Ns
                ins = np.array([v[t][b + k * nb] * np.exp(sign * 2j * np.pi * a * k / (Ns * r)) for k in range(r)])
                y = np.array([sum(ins[q] * np.exp(sign * 2j * np.pi * q * k / r) for q in range(r)) for k in range(r)])
                base = (j // Ns) * Ns * r + a
                for k in range(r):
                    idx = base + k * Ns
                    if last:
                        assert idx == t + (b + k * nb) * T, "last pass must land on register positions"
                    out[idx] = y[k]
        Ns *= r
        v = np.array([[out[t + m * T] for m in range(R)] for t in range(T)], dtype=complex)
    res = np.zeros(N, dtype=complex)
    for t in range(T):
        for m in range(R):
            res[t + m * T] = v[t][m]
    return res


def r2c_post(Z):
    """X[k], k=0..N from Z = FFT_N(x[0::2] + i x[1::2]); forward sign -1."""
    N = len(Z)
    k = np.arange(N + 1)
    Zk = Z[k % N]
    Zc = np.conj(Z[(N - k) % N])
    w = np.exp(-2j * np.pi * k / (2 * N))
    return 0.5 * ((Zk + Zc) - 1j * w * (Zk - Zc))


def c2r_pre(X):
    """Z[k], k=0..N-1 such that ifft_N(Z) (unnormalised, sign +1) = (x[0::2] + i x[1::2]) * N ... see check()."""
    N = len(X) - 1
    k = np.arange(N)
    Xk = X[k]
    Xc = np.conj(X[N - k])
    w = np.exp(+2j * np.pi * k / (2 * N))
    return (Xk + Xc) + 1j * w * (Xk - Xc)


def bank_conflicts(N, R, radices, wordbytes, pad_every, pad_words=1, W=1):
    """Worst conflict degree of the scattered store of each non-final pass.  Element = one real of `wordbytes`
    (re/im split exchange).  A warp = 32 consecutive threads; a wavefront covers 128 bytes worth of lanes
    (32 lanes for 4-byte words, 16 for 8-byte words, 8 for the 16-byte complex words Float64 uses since round 2)."""
    T = N // R
    lanes = 128 // wordbytes          # a wavefront covers 128 bytes worth of lanes: 32 / 16 / 8 lanes for 4- / 8- / 16-byte words
    res = []
    Ns = 1
    for p, r in enumerate(radices[:-1]):
        nb = R // r
        worst = 1
        for b in range(nb):
            for k in range(r):
                for t0 in range(0, T, lanes):
                    banks = {}
                    for t in range(t0, min(t0 + lanes, T)):
                        j = t + b * T
                        idx = (j // Ns) * Ns * r + (j % Ns) + k * Ns
                        phys = idx + (idx // pad_every) * pad_words
                        bank = (phys * (wordbytes // 4)) % 32      # first 4-byte bank of the word; words are aligned to their size
                        banks[bank] = banks.get(bank, 0) + 1
                    worst = max(worst, max(banks.values()))
        res.append(worst)
        Ns *= r
    # read side: idx = t + m*T
    worst = 1
    for m in range(R):
        for t0 in range(0, T, lanes):
            banks = {}
            for t in range(t0, min(t0 + lanes, T)):
                idx = t + m * T
                phys = idx + (idx // pad_every) * pad_words
                bank = (phys * (wordbytes // 4)) % 32      # first 4-byte bank of the word; words are aligned to their size
                banks[bank] = banks.get(bank, 0) + 1
            worst = max(worst, max(banks.values()))
    return res, worst


def c3_step_traffic(n, aliased_fraction=1 / 3, es=8):
    """DRAM bytes one ETDRK4 step of the fused 2-D vorticity problem (config C3) moves in THIS implementation, kernel class by kernel
    class, for an n x n grid whose y dimension is four-step (n >= 4096).  S = spectral array, P = physical array, R = dense real
    spectral-shaped array; f / g = live fraction of the kx columns / ky rows (alias ranges: src/domains.jl:408-427).
      inverse (zeta, u, v from one sol, ffb_fft_inverse_multi): shared sub-pass A reads S + R (invKrsq once) and writes 3 S;
        3 sub-passes B: 2 S each; 3 c2r row passes: S -> P, two of them also read the zeta field.
      forward 1 (dealias = 2, box don't-care): r2c P -> f S; sub-pass A f S -> f S; sub-pass B f S -> f g S.
      forward 2 (dealias = 1): r2c P -> f S; sub-pass A f S -> f S; sub-pass B reads f S + f g S (accumulated), writes S (zeros included).
      stages (stages.cu, Float64 dense ETD coefficients): 2 x (3 S + 2 R) + (4 S + 2 R) + (6 S + 4 R)."""
    nkr = n // 2 + 1
    S, P, R = nkr * n * 2 * es, n * n * es, nkr * n * es
    iL = int(np.floor((1 - aliased_fraction) / 2 * n)) + 1
    iR = int(np.ceil((1 + aliased_fraction) / 2 * n))
    f = 1 - (nkr - iL + 1) / nkr if aliased_fraction > 0 else 1.0
    g = 1 - (iR - iL + 1) / n if aliased_fraction > 0 else 1.0
    per_calcN = {
        "fs_am (shared inverse sub-pass A)": S + R + 3 * S,
        "fs_b inverse": 3 * 2 * S,
        "c2r rows": 3 * (S + P) + 2 * P,
        "r2c rows": 2 * (P + f * S),
        "fs_a forward": 2 * 2 * f * S,
        "fs_b forward": (f * S + f * g * S) + (f * S + f * g * S + S),
    }
    out = {k: 4 * v for k, v in per_calcN.items()}
    out["stages"] = 2 * (3 * S + 2 * R) + (4 * S + 2 * R) + (6 * S + 4 * R)
    out["total"] = sum(out.values())
    return out


def check():
    rng = np.random.default_rng(0)
    for N, R, rad in [(64, 8, [8, 8]), (128, 16, [16, 8]), (256, 16, [16, 16]), (512, 8, [8, 8, 8]), (512, 16, [16, 16, 2]),
                      (1024, 16, [16, 16, 4]), (32, 8, [8, 4]), (16, 16, [16]), (64, 16, [16, 4]), (2048, 16, [16, 16, 8])]:
        x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        for sign in (-1, 1):
            y = stockham(x, rad, R, sign)
            ref = np.fft.fft(x) if sign < 0 else np.fft.ifft(x) * N
            err = np.linalg.norm(y - ref) / np.linalg.norm(ref)
            assert err < 1e-12, (N, R, rad, sign, err)
    print("stockham recurrence OK")
    for n in (8, 64, 30):
        x = rng.standard_normal(n)
        z = x[0::2] + 1j * x[1::2]
        X = r2c_post(np.fft.fft(z))
        assert np.allclose(X, np.fft.rfft(x), atol=1e-12)
        Z = c2r_pre(np.fft.rfft(x))
        zz = np.fft.ifft(Z) * (n // 2)  # unnormalised inverse of length N = n/2
        xr = np.empty(n)
        xr[0::2], xr[1::2] = zz.real, zz.imag
        assert np.allclose(xr / n, x, atol=1e-12), "c2r: unnormalised inverse of c2r_pre(X) = n * x"
    print("r2c/c2r formulas OK  (c2r: x = interleave(ifft_unnorm(c2r_pre(X))) / nx)")
    for wb in (8, 4):
        for N, R in [(8192, 16), (4096, 16), (2048, 16), (1024, 16), (512, 16), (256, 16), (128, 16), (64, 16), (4096, 8), (512, 8), (64, 8)]:
            rad = radix_plan(N, R)
            for pad in ((16, 1) if wb == 8 else (32, 1), (32, 2) if wb == 8 else (32, 1)):
                w, rd = bank_conflicts(N, R, rad, wb, pad[0], pad[1])
                print(f"word {wb}B N={N:5d} R={R:2d} radices={rad} pad +{pad[1]}/{pad[0]}: store conflicts {w}, read {rd}")


if __name__ == "__main__":
    check()
