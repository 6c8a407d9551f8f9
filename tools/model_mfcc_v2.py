#!/usr/bin/env python3
"""numpy model of the index flow of k_mfcc_fused v2 (real 2048-point FFT as 64-real x 32-complex, no cross-lane
post-pass, special columns k1 = 0 / 32 in a helper warp).  Validates every index map used by the CUDA code."""
import numpy as np

def br5(k):
    return int("{:05b}".format(k)[::-1], 2)

def model(frame_windowed_half):
    """frame_windowed_half: x * window * 0.5 (2048).  Returns power spectrum P[0..1024] through the v2 data flow."""
    xh = frame_windowed_half.astype(np.float64)
    # stage 1: lane n2 holds s[n1] = xh[32 n1 + n2]; z[m] = s[2m] + i s[2m+1]; Z = DFT32(z)
    c = np.zeros((32, 32), complex)          # c[n2][k1]  slot 0 = (R0, R32) packed as complex
    for n2 in range(32):
        s = xh[32 * np.arange(64) + n2]
        z = s[0::2] + 1j * s[1::2]
        Z = np.fft.fft(z)
        R = np.zeros(33, complex)
        for k in range(1, 16):
            Zk, Zm = Z[k], Z[32 - k]
            E = Zk + np.conj(Zm)
            O = Zk - np.conj(Zm)
            T = O * (-1j) * np.exp(-2j * np.pi * k / 64)
            R[k] = E + T
            R[32 - k] = np.conj(E - T)
        R[16] = 2 * np.conj(Z[16])
        R[0] = 2 * (Z[0].real + Z[0].imag)
        R[32] = 2 * (Z[0].real - Z[0].imag)
        # check: R == DFT64 of 2*s (window carried the 1/2)
        ref = np.fft.fft(2 * s)[:33]
        assert np.allclose(R, ref, atol=1e-9), (n2, np.abs(R - ref).max())
        c[n2, 0] = R[0].real + 1j * R[32].real
        c[n2, 1:32] = R[1:32]
    special = c[:, 0].copy()                 # special[n2] = (a, b)
    # twiddle W_2048^(n2 k1), transposition, stage 2 in lane k1 = 1..31
    P = np.zeros(1025)
    for k1 in range(1, 32):
        y = c[:, k1] * np.exp(-2j * np.pi * np.arange(32) * k1 / 2048)
        Y = np.fft.fft(y)
        for k2 in range(32):
            b = k1 + 64 * k2 if k2 < 16 else 64 * (32 - k2) - k1
            P[b] = abs(Y[k2]) ** 2
    # special pass: kind 0: DFT32(a) -> bins 64 k2 (k2 = 0..16); kind 1: DFT32(b W_64^n2) -> bins 32 + 64 k2 (k2 = 0..15)
    a, b = special.real, special.imag
    A = np.fft.fft(a)
    Cc = np.fft.fft(b * np.exp(-2j * np.pi * np.arange(32) / 64))
    for k2 in range(17):
        P[64 * k2] = abs(A[k2]) ** 2
    for k2 in range(16):
        P[32 + 64 * k2] = abs(Cc[k2]) ** 2
    return P

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    x = rng.standard_normal(2048)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(2048) / 2048)
    P = model(x * w * 0.5)
    ref = np.abs(np.fft.rfft(x * w)) ** 2
    print("max rel err", np.abs(P - ref).max() / ref.max())
    assert np.allclose(P, ref, rtol=1e-9, atol=1e-9 * ref.max())
