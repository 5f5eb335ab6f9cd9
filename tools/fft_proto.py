"""Numpy prototype of the ring-FFT decomposition used by calclens_b200/csrc/ring_fft.cu (index math check).

 * real ring of length n = 4r  ->  two complex DFTs of length r (radix-4 split, two real sequences per complex one)
 * DFT_r for arbitrary r: Bluestein with power-of-two M >= 2r-1, in-place DIF forward (bit-reversed spectrum),
   pointwise product with a bit-reversed chirp spectrum, in-place DIT inverse (natural order out)
Run: python tools/fft_proto.py
"""
import numpy as np


def dif_forward(a):
    """in-place radix-2 decimation-in-frequency, natural in -> bit-reversed out, sign -1"""
    M = a.size
    span = M // 2
    while span >= 1:
        for base in range(0, M, 2 * span):
            j = np.arange(span)
            w = np.exp(-2j * np.pi * j / (2 * span))
            u = a[base + j].copy(); v = a[base + span + j].copy()
            a[base + j] = u + v
            a[base + span + j] = (u - v) * w
        span //= 2
    return a


def dit_inverse(a):
    """in-place radix-2 decimation-in-time, bit-reversed in -> natural out, sign +1 (unnormalised)"""
    M = a.size
    span = 1
    while span < M:
        for base in range(0, M, 2 * span):
            j = np.arange(span)
            w = np.exp(+2j * np.pi * j / (2 * span))
            u = a[base + j].copy(); v = a[base + span + j].copy() * w
            a[base + j] = u + v
            a[base + span + j] = u - v
        span *= 2
    return a


def bluestein_tables(r):
    M = 1
    while M < 2 * r - 1:
        M *= 2
    j = np.arange(r)
    chirp = np.exp(-1j * np.pi * ((j * j) % (2 * r)) / r)      # w_j = exp(-i pi j^2 / r)
    b = np.zeros(M, complex)
    b[:r] = np.conj(chirp)
    b[M - np.arange(1, r)] = np.conj(chirp[1:])
    bhat_br = dif_forward(b.copy())                               # bit-reversed order, as the kernel stores it
    return M, chirp, bhat_br


def dft_r(z, tabs=None):
    """forward DFT of arbitrary length r through the kernel's pipeline"""
    r = z.size
    if r & (r - 1) == 0:
        a = dif_forward(z.astype(complex).copy())
        # bit reversal on read-out
        bits = r.bit_length() - 1
        idx = np.array([int(format(i, "0%db" % bits)[::-1], 2) if bits else 0 for i in range(r)])
        return a[idx]
    M, chirp, bhat_br = tabs if tabs is not None else bluestein_tables(r)
    a = np.zeros(M, complex)
    a[:r] = z * chirp
    dif_forward(a)
    a *= bhat_br
    dit_inverse(a)
    return a[:r] / M * chirp


def ring_analysis(x):
    """r2c of a real ring of length n=4r: F_k, k = 0..n/2"""
    n = x.size; r = n // 4
    z1 = x[0::4] + 1j * x[1::4]
    z2 = x[2::4] + 1j * x[3::4]
    tabs = None if r & (r - 1) == 0 else bluestein_tables(r)
    Z1 = dft_r(z1, tabs); Z2 = dft_r(z2, tabs)
    k = np.arange(n // 2 + 1)
    kk = k % r
    kc = (r - kk) % r
    X0 = 0.5 * (Z1[kk] + np.conj(Z1[kc])); X1 = -0.5j * (Z1[kk] - np.conj(Z1[kc]))
    X2 = 0.5 * (Z2[kk] + np.conj(Z2[kc])); X3 = -0.5j * (Z2[kk] - np.conj(Z2[kc]))
    W = np.exp(-2j * np.pi * k / n)
    return X0 + W * X1 + W ** 2 * X2 + W ** 3 * X3


def ring_synthesis(Y):
    """c2r (unnormalised) from the half spectrum Y_k, k = 0..n/2, n = 4r"""
    n = 2 * (Y.size - 1); r = n // 4
    Yf = np.zeros(n, complex)
    Yf[: n // 2 + 1] = Y
    Yf[0] = Y[0].real; Yf[n // 2] = Y[n // 2].real
    Yf[n // 2 + 1:] = np.conj(Yf[1: n // 2][::-1])
    kp = np.arange(r)
    U = []
    for q in range(4):
        u = np.zeros(r, complex)
        for p in range(4):
            u += Yf[kp + p * r] * np.exp(2j * np.pi * q * (kp + p * r) / n)
        U.append(u)
    V1 = U[0] + 1j * U[1]; V2 = U[2] + 1j * U[3]
    tabs = None if r & (r - 1) == 0 else bluestein_tables(r)
    x01 = np.conj(dft_r(np.conj(V1), tabs)); x23 = np.conj(dft_r(np.conj(V2), tabs))
    x = np.zeros(n)
    x[0::4] = x01.real; x[1::4] = x01.imag; x[2::4] = x23.real; x[3::4] = x23.imag
    return x


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for r in [1, 2, 3, 5, 8, 13, 64, 100, 255, 256]:
        n = 4 * r
        x = rng.normal(size=n)
        F = ring_analysis(x)
        Fr = np.fft.rfft(x)
        e1 = np.abs(F - Fr).max() / np.abs(Fr).max()
        Y = rng.normal(size=n // 2 + 1) + 1j * rng.normal(size=n // 2 + 1)
        xs = ring_synthesis(Y)
        Y2 = Y.copy(); Y2[0] = Y2[0].real; Y2[-1] = Y2[-1].real
        xr = np.fft.irfft(Y2, n) * n
        e2 = np.abs(xs - xr).max() / np.abs(xr).max()
        print("r=%4d  analysis err %.2e  synthesis err %.2e" % (r, e1, e2))
        assert e1 < 1e-12 and e2 < 1e-12
