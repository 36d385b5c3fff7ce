"""NumPy model of the CUDA kernels' index algebra (development aid; not shipped, not used by tests).
Mirrors: Stockham pass structure, DIF butterflies with bit-reversed register order, paired last
pass + Hermitian split of the row kernel, two-level column FFT with inter-level twiddle."""
import numpy as np

def bitrev(i, bits):
    r = 0
    for b in range(bits):
        r |= ((i >> b) & 1) << (bits - 1 - b)
    return r

def fft_dif(v):
    """in-place radix-2 DIF on list of R complex; output bit-reversed (v[i] = X[bitrev(i)])"""
    R = len(v)
    h = R // 2
    while h >= 1:
        for b in range(0, R, 2 * h):
            for q in range(h):
                a, c = v[b + q], v[b + q + h]
                v[b + q] = a + c
                v[b + q + h] = (a - c) * np.exp(-2j * np.pi * q / (2 * h))
        h //= 2
    return v

def stockham_pass(x, n, R, Ns):
    T = n // R
    bits = R.bit_length() - 1
    out = np.zeros(n, complex)
    for j in range(T):
        k = j % Ns
        v = [x[j + r * T] * np.exp(-2j * np.pi * r * k / (Ns * R)) for r in range(R)]
        fft_dif(v)
        j0 = (j // Ns) * Ns * R + k
        for s in range(R):
            out[j0 + s * Ns] = v[bitrev(s, bits)]
    return out

def fft_stockham(x, radices):
    n = len(x); Ns = 1
    for R in radices:
        x = stockham_pass(x, n, R, Ns); Ns *= R
    assert Ns == n
    return x

def r2c_row_model(xr, prefix, RL=16):
    """row kernel: m = prod(prefix)*RL; last pass paired + Hermitian split"""
    n = len(xr); m = n // 2
    z = xr[0::2] + 1j * xr[1::2]
    Ns = 1
    for R in prefix:
        z = stockham_pass(z, m, R, Ns); Ns *= R
    PP = Ns; assert PP * RL == m
    bits = RL.bit_length() - 1
    X = np.zeros(m + 1, complex)
    wn = lambda k: np.exp(-2j * np.pi * k / n)
    def bfly(j):
        v = [z[j + r * PP] * np.exp(-2j * np.pi * r * j / m) for r in range(RL)]
        fft_dif(v)
        return [v[bitrev(s, bits)] for s in range(RL)]  # natural order: Z[j + s*PP]
    def pair(a, b, k):
        # a = Z[k], b = Z[m-k] -> X[k], X[m-k]
        s = a + np.conj(b); d = a - np.conj(b)
        t = (0.5 * wn(k)) * d
        E = 0.5 * s
        return (E.real + t.imag) + 1j * (E.imag - t.real), (E.real - t.imag) - 1j * (E.imag + t.real)
    for lt in range(max(PP // 2, 1)):
        if PP == 1:
            raise NotImplementedError
        jA = lt; jB = PP // 2 if lt == 0 else PP - lt
        A = bfly(jA); B = bfly(jB)
        if lt > 0:
            for s in range(RL):
                kA = jA + s * PP
                X[kA], X[m - kA] = pair(A[s], B[RL - 1 - s], kA)
        else:
            X[0] = A[0].real + A[0].imag; X[m] = A[0].real - A[0].imag
            for s in range(1, RL // 2):
                k = s * PP
                X[k], X[m - k] = pair(A[s], A[RL - s], k)
            k = (RL // 2) * PP
            X[k] = np.conj(A[RL // 2])
            for s in range(RL // 2):
                k = PP // 2 + s * PP
                X[k], X[m - k] = pair(B[s], B[RL - 1 - s], k)
    return X

def col_two_level(x, n1, n2, pl1, pl2):
    """x[j1*n2 + j2] -> X[k1 + n1*k2]"""
    n = n1 * n2
    a = x.reshape(n1, n2)
    S = np.zeros((n1, n2), complex)
    for j2 in range(n2):  # A tiles
        col = fft_stockham(a[:, j2].copy(), pl1)
        S[:, j2] = col * np.exp(-2j * np.pi * np.arange(n1) * j2 / n)
    out = np.zeros(n, complex)
    for k1 in range(n1):  # B tiles
        row = fft_stockham(S[k1].copy(), pl2)
        out[k1 + n1 * np.arange(n2)] = row
    return out

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for n, pl in [(16, [16]), (32, [8, 4]), (64, [8, 8]), (128, [16, 8]), (512, [8, 8, 8]), (32, [32]), (8, [2, 4])]:
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        print("c2c", n, pl, np.abs(fft_stockham(x, pl) - np.fft.fft(x)).max())
    for m, prefix in [(32, [2]), (64, [4]), (512, [32]), (1024, [8, 8]), (2048, [16, 8])]:
        xr = rng.standard_normal(2 * m)
        print("r2c", m, prefix, np.abs(r2c_row_model(xr, prefix) - np.fft.rfft(xr)).max())
    for n1, n2, p1, p2 in [(16, 16, [16], [16]), (32, 16, [8, 4], [16]), (8, 64, [8], [8, 8])]:
        x = rng.standard_normal(n1 * n2) + 1j * rng.standard_normal(n1 * n2)
        print("2lvl", n1, n2, np.abs(col_two_level(x, n1, n2, p1, p2) - np.fft.fft(x)).max())
