"""Index model of the cluster (DSMEM) four-step kernel csrc/fft_cluster.cuh: checks that the thread->element maps of
step A, the remote scatter and step B reproduce an N-point DFT, and counts bank conflicts of the receive buffer."""
import numpy as np


def model(N1, N2, C, W, sign=-1, seed=0):
    N = N1 * N2
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((N, W)) + 1j * rng.standard_normal((N, W))      # tile: N rows x W columns
    ref = np.fft.fft(x, axis=0) if sign < 0 else np.fft.ifft(x, axis=0) * N
    Tn1, Tn2 = N1 // 16, N2 // 16
    NLA, NLB = (N2 // C) * W, (N1 // C) * W
    NT = NLA * Tn1
    assert NT == NLB * Tn2, (NT, NLB * Tn2)
    pad = 0 if W * 16 >= 128 else (64 // 16)
    RS = (N1 // C) * W + pad
    rbuf = [np.zeros(N2 * RS, dtype=complex) for _ in range(C)]
    # ---- step A + scatter
    for c in range(C):
        for tid in range(NT):
            lw, t = tid % NLA, tid // NLA
            w, n2l = lw % W, lw // W
            n2 = c * (N2 // C) + n2l
            v = np.array([x[N2 * (t + m * Tn1) + n2, w] for m in range(16)])
            # sub-FFT over n1: result positions k1 = t + m*Tn1 (Stockham lands on register positions) -> emulate with a DFT of the sub-line
            sub = np.array([x[N2 * n1 + n2, w] for n1 in range(N1)])
            Y = np.fft.fft(sub) if sign < 0 else np.fft.ifft(sub) * N1
            for m in range(16):
                k1 = t + m * Tn1
                val = Y[k1] * np.exp(sign * 2j * np.pi * ((n2 * k1) % N) / N)
                d, k1l = k1 // (N1 // C), k1 % (N1 // C)
                rbuf[d][n2 * RS + k1l * W + w] = val
    out = np.zeros((N, W), dtype=complex)
    # ---- step B
    for c in range(C):
        for tid in range(NT):
            lw, t2 = tid % NLB, tid // NLB
            w, k1l = lw % W, lw // W
            sub = np.array([rbuf[c][n2 * RS + k1l * W + w] for n2 in range(N2)])
            Z = np.fft.fft(sub) if sign < 0 else np.fft.ifft(sub) * N2
            k1 = c * (N1 // C) + k1l
            for m in range(16):
                k2 = t2 + m * Tn2
                out[k1 + N1 * k2, w] = Z[k2]
    err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    # ---- bank conflicts on the remote write (16-byte elements: 8 lanes per wavefront, 8 bank groups of 16 B)
    worst_w = 1
    for tid0 in range(0, NT, 8):
        for m in range(16):
            banks = {}
            for tid in range(tid0, tid0 + 8):
                lw, t = tid % NLA, tid // NLA
                w, n2l = lw % W, lw // W
                k1 = t + m * Tn1
                idx = n2l * RS + (k1 % (N1 // C)) * W + w
                dest = k1 // (N1 // C)
                b = (dest, idx % 8)
                banks[b] = banks.get(b, 0) + 1
            worst_w = max(worst_w, max(banks.values()))
    worst_r = 1
    for tid0 in range(0, NT, 8):
        for m in range(16):
            banks = {}
            for tid in range(tid0, tid0 + 8):
                lw, t2 = tid % NLB, tid // NLB
                w, k1l = lw % W, lw // W
                idx = (t2 + m * Tn2) * RS + k1l * W + w
                banks[idx % 8] = banks.get(idx % 8, 0) + 1
            worst_r = max(worst_r, max(banks.values()))
    return err, NT, RS, worst_w, worst_r


if __name__ == "__main__":
    for cfg in [(64, 128, 8, 4), (64, 64, 8, 4), (64, 128, 16, 8), (32, 64, 8, 8), (64, 128, 8, 8)]:
        for sign in (-1, 1):
            err, nt, rs, ww, wr = model(*cfg, sign=sign)
            print(cfg, "sign", sign, f"err {err:.1e} threads {nt} RS {rs} smem(rbuf, 16B elems) {cfg[1] * rs * 16 / 1024:.1f} KB write-conflict {ww} read-conflict {wr}")
