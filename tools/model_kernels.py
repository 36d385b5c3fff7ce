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


# ---------------------------------------------------------------------------------------------------------------------
# round 2 kernels
# ---------------------------------------------------------------------------------------------------------------------
def w(n, e):
    return np.exp(-2j * np.pi * (e % n) / n)


def swizzle(p):
    """kernels_rows_v2.cuh: position of element p inside the shared-memory pencil (16-byte units)."""
    return p ^ (((p >> 4) ^ (p >> 9)) & 7)


def rows_v2_model(z, nsub=16):
    """kernels_rows_v2.cuh: m = nsub x (m/nsub); warp-local in-place DIF (radix 16 then radix q/16) of the stride-nsub
    sub-sequences in a swizzled pencil, then the radix-nsub pass across the sub-sequences.  Returns Z = FFT(z)."""
    m = len(z)
    q = m // nsub                      # 512 for m = 8192
    ra, rb = 16, q // 16               # pass A radix 16 over stride rb, pass B radix rb
    sm = np.zeros(m + 8, complex)
    for p in range(m):
        sm[swizzle(p)] = z[p]
    for j1 in range(nsub):             # pass A
        for u in range(rb):
            pos = [swizzle(j1 + nsub * (u + rb * r)) for r in range(ra)]
            a = sm[pos]
            b = np.array([sum(w(ra, r * s) * a[r] for r in range(ra)) * w(q, u * s) for s in range(ra)])
            sm[pos] = b
    for j1 in range(nsub):             # pass B
        for s in range(ra):
            pos = [swizzle(j1 + nsub * (u + rb * s)) for u in range(rb)]
            c = sm[pos]
            sm[pos] = np.array([sum(w(rb, u * t) * c[u] for u in range(rb)) for t in range(rb)])
    Z = np.zeros(m, complex)
    for k2 in range(q):                # final pass: F_j1[k2] at j1 + nsub*(t + rb*s), k2 = s + 16 t
        s, t = k2 % ra, k2 // ra
        f = np.array([sm[swizzle(j1 + nsub * (t + rb * s))] * w(m, j1 * k2) for j1 in range(nsub)])
        for k1 in range(nsub):
            Z[k2 + q * k1] = sum(w(nsub, j1 * k1) * f[j1] for j1 in range(nsub))
    return Z


def herm_split(Z, n):
    """X[k] = E[k] - i w_n^k O[k], k = 0..m, from the half-length spectrum Z (m = n/2)."""
    m = n // 2
    X = np.zeros(m + 1, complex)
    for k in range(m + 1):
        a, b = Z[k % m], np.conj(Z[(m - k) % m])
        X[k] = 0.5 * (a + b) - 0.5j * w(n, k) * (a - b)
    return X


def rows_long2_model(x):
    """kernels_rows_long2.cuh: two DIF halves of a long row; the even half parks X[2 k2], the odd half produces X[2 k2 + 1] for
    the same k2 and both leave as one pair.  Returns the r2c spectrum X[0..m]."""
    n = len(x)
    m = n // 2
    M = m // 2
    z = x[0::2] + 1j * x[1::2]
    y0 = z[:M] + z[M:]
    y1 = (z[:M] - z[M:]) * np.array([w(m, j) for j in range(M)])
    Z0, Z1 = np.fft.fft(y0), np.fft.fft(y1)           # Z[2 k2], Z[2 k2 + 1]
    out = np.zeros(m + 1, complex)
    parked = np.zeros(M + 1, complex)
    for k2 in range(M + 1):                           # even half: partner of k2 is M - k2
        a, b = Z0[k2 % M], np.conj(Z0[(M - k2) % M])
        parked[k2] = 0.5 * (a + b) - 0.5j * w(n, 2 * k2) * (a - b)
    for k2 in range(M):                               # odd half: partner of k2 is M - 1 - k2; emit the pair (2 k2, 2 k2 + 1)
        a, b = Z1[k2], np.conj(Z1[M - 1 - k2])
        out[2 * k2] = parked[k2]
        out[2 * k2 + 1] = 0.5 * (a + b) - 0.5j * w(n, 2 * k2 + 1) * (a - b)
    out[m] = parked[M]
    return out


def herm_pair(a, b, wv):
    """kernels_rows.cuh: herm_pair(Z[k], Z[m-k], w_n^k) -> X[k], X[m-k]."""
    s, d = a + np.conj(b), a - np.conj(b)
    t = d * (0.5 * wv)
    e = 0.5 * s
    xk = complex(e.real + t.imag, e.imag - t.real)
    xmk = complex(e.real - t.imag, -(e.imag + t.real))
    return xk, xmk


def rows_dit2_model(x, PP):
    """kernels_rows_dit2.cuh: n = 4 M real points, M = 16 PP.  Decimation in time: Ze = FFT_M(z[2j]) is parked per thread, the
    second half computes Zo = FFT_M(z[2j+1]) with the same thread <-> column mapping and every thread combines
    {Ze, Zo}[k], {Ze, Zo}[M-k] into X[k], X[M-k], X[k+M], X[2M-k].  Returns (X, number of stores per bin)."""
    n = len(x)
    M = n // 4
    assert M == 16 * PP
    m = 2 * M
    z = x[0::2] + 1j * x[1::2]
    Ze, Zo = np.fft.fft(z[0::2]), np.fft.fft(z[1::2])
    X = np.zeros(m + 1, complex)
    cnt = np.zeros(m + 1, int)

    def put(k, v):
        X[k] = v
        cnt[k] += 1

    def four(k, ze_k, zo_k, ze_mk, zo_mk, second=True):
        wm, wn = w(m, k), w(n, k)
        P, Q = zo_k * wm, -(zo_mk * np.conj(wm))
        xk, x2mk = herm_pair(ze_k + P, ze_mk - Q, wn)
        put(k, xk)
        put(2 * M - k, x2mk)
        if second:
            xkM, xMk = herm_pair(ze_k - P, ze_mk + Q, wn * -1j)
            put(k + M, xkM)
            put(M - k, xMk)

    for lt in range(PP // 2):
        jA = lt
        if lt:
            jB = PP - lt
            for s in range(16):
                k, mk_ = jA + s * PP, jB + (15 - s) * PP
                assert k + mk_ == M
                four(k, Ze[k], Zo[k], Ze[mk_], Zo[mk_])
        else:
            z0 = Ze[0] + Zo[0]                       # Z[0]; Z[M] = Ze[0] - Zo[0]
            put(0, z0.real + z0.imag)
            put(m, z0.real - z0.imag)
            put(M, np.conj(Ze[0] - Zo[0]))
            for s in range(1, 8):
                k, mk_ = s * PP, (16 - s) * PP
                four(k, Ze[k], Zo[k], Ze[mk_], Zo[mk_])
            k = 8 * PP                               # M/2 is its own partner: X[M/2], X[3M/2]
            four(k, Ze[k], Zo[k], Ze[k], Zo[k], second=False)
            for s in range(8):
                k, mk_ = PP // 2 + s * PP, PP // 2 + (15 - s) * PP
                four(k, Ze[k], Zo[k], Ze[mk_], Zo[mk_])
    return X, cnt


def rows_ditc_model(x, C, PP):
    """kernels_rows_ditc.cuh: n = 2 C M real points, M = 16 PP.  C classes by sample index mod C, Zc = FFT_M(z[C j + c]) parked per
    thread; the combine forms, for a pair (k, M - k),  u_c = w_m^(c k) Zc[k],  v_c = conj(w_m^(c k)) Zc[M - k],  U = DFT_C(u),
    V = DFT_C(v)  and  X[k + M q], X[(M - k) + M (C - 1 - q)] = herm_pair(U[q], V[(C - q) mod C], w_n^(k + M q)).
    Returns (X, number of stores per bin)."""
    n = len(x)
    M = n // (2 * C)
    assert M == 16 * PP
    m = C * M
    z = x[0::2] + 1j * x[1::2]
    Zc = [np.fft.fft(z[c::C]) for c in range(C)]
    X = np.zeros(m + 1, complex)
    cnt = np.zeros(m + 1, int)

    def put(k, v):
        X[k] = v
        cnt[k] += 1

    tw = [w(n, i) for i in range(n)]                       # the plan's table
    W = [tw[i * PP] for i in range(32 * C)]                # shared-memory table w_{32C}^i

    def pair(k, mk_, nq, last_single=False):
        # k and mk_ index the sub-spectra (mk_ = (M - k) mod M); the kernel factors the twiddles over k = j + s PP
        j, s = k % PP, k // PP
        wck = [tw[2 * c * j] * W[(2 * c * s) % (32 * C)] for c in range(C)]     # w_m^(c k)
        assert all(abs(wck[c] - w(m, c * k)) < 1e-12 for c in range(C))
        u = np.array([wck[c] * Zc[c][k] for c in range(C)])
        v = np.array([np.conj(wck[c]) * Zc[c][mk_] for c in range(C)])
        U = np.array([sum(w(C, c * q) * u[c] for c in range(C)) for q in range(C)])
        Vv = np.array([sum(w(C, c * q) * v[c] for c in range(C)) for q in range(C)])
        for q in range(nq):
            wn = tw[j] * W[s] * W[16 * q]                                        # w_n^(k + M q)
            assert abs(wn - w(n, k + M * q)) < 1e-12
            xk, xmk = herm_pair(U[q], Vv[(C - q) % C], wn)
            put(k + M * q, xk)
            if not (last_single and q == nq - 1):
                put((M - k) + M * (C - 1 - q), xmk)

    for lt in range(PP // 2):
        if lt:
            for s in range(16):
                k = lt + s * PP
                pair(k, M - k, C)
        else:
            pair(0, 0, C // 2 + 1, last_single=True)      # bins M q: q = 0 gives X[0] and X[m]; q = C/2 is its own partner
            for s in range(1, 8):
                pair(s * PP, (16 - s) * PP, C)
            pair(8 * PP, 8 * PP, C // 2)                   # k = M/2 is its own partner
            for s in range(8):
                pair(PP // 2 + s * PP, PP // 2 + (15 - s) * PP, C)
    return X, cnt


def gen_rev(k, q, lg):
    """kernels_generic.cuh: position of F[k] after the in-place DIF passes (radix 4 ..., one radix 2 when lg is odd)."""
    pos, length = 0, q
    for _ in range(lg // 2):
        length >>= 2
        pos += (k & 3) * length
        k >>= 2
    if lg & 1:
        pos += k & 1
    return pos


def rows_mixed_model(z, t):
    """kernels_generic.cuh: m = t * q; in-place radix-4 DIF FFT_q of the stride-t sub-sequences, radix-t combine."""
    m = len(z)
    q = m // t
    lg = q.bit_length() - 1
    sm = np.array(z, complex)
    length = q
    for _ in range(lg // 2):
        quarter = length // 4
        for j1 in range(t):
            for b in range(q // 4):
                blk, i = divmod(b, quarter)
                pos = [j1 + t * (blk * length + i + r * quarter) for r in range(4)]
                a = sm[pos]
                y = np.array([sum(w(4, r * s) * a[r] for r in range(4)) * w(length, i * s) for s in range(4)])
                sm[pos] = y
        length = quarter
    if lg & 1:
        for j1 in range(t):
            for b in range(q // 2):
                p0, p1 = j1 + t * 2 * b, j1 + t * (2 * b + 1)
                a0, a1 = sm[p0], sm[p1]
                sm[p0], sm[p1] = a0 + a1, a0 - a1
    Z = np.zeros(m, complex)
    for k2 in range(q):
        slot = t * gen_rev(k2, q, lg)
        f = np.array([sm[slot + j1] * w(m, j1 * k2) for j1 in range(t)])
        for k1 in range(t):
            Z[k2 + q * k1] = sum(w(t, j1 * k1) * f[j1] for j1 in range(t))
    return Z


def cols_mixed_model(y, t):
    """kernels_generic.cuh: nx = t * q; odd-radix pre-stage A[k1][x2] = w_nx^(k1 x2) sum_x1 w_t^(x1 k1) y[x1 q + x2], then a
    length-q transform per k1 whose output k2 is row k1 + t k2 (ColDst::vt)."""
    nx = len(y)
    q = nx // t
    out = np.zeros(nx, complex)
    for k1 in range(t):
        a = np.array([w(nx, k1 * x2) * sum(w(t, x1 * k1) * y[x1 * q + x2] for x1 in range(t)) for x2 in range(q)])
        out[k1 + t * np.arange(q)] = np.fft.fft(a)
    return out


def cols_split2_model(y, n1, n2):
    """kernels_cols.cuh, SPLIT = 2: radix-2 DIF pre-stage folded into the level-A load, rows kx = c2 + 2 k'."""
    nx = len(y)
    h = nx // 2
    out = np.zeros(nx, complex)
    for c2 in range(2):
        yc = (y[:h] + y[h:]) if c2 == 0 else (y[:h] - y[h:]) * np.array([w(nx, x) for x in range(h)])
        out[c2 + 2 * np.arange(h)] = np.fft.fft(yc)
    return out


def bluestein_model(x):
    """kernels_bluestein.cuh / launch_bluestein.cu: X = c .* IFFT_M(FFT_M(x .* c, padded) .* FFT_M(h)), c[j] = exp(-i pi j^2/n),
    h[l] = conj(c[|l|]) wrapped to length M = 2^p >= 2n-1; the inverse transform as conj(FFT(conj(.)))/M."""
    n = len(x)
    M = 1
    while M < 2 * n - 1:
        M <<= 1
    j = np.arange(n)
    c = np.exp(-1j * np.pi * ((j * j) % (2 * n)) / n)
    h = np.zeros(M, complex)
    h[:n] = np.conj(c)
    h[M - np.arange(1, n)] = np.conj(c[1:])
    a = np.zeros(M, complex)
    a[:n] = x * c
    P = np.fft.fft(a) * np.fft.fft(h)
    p = np.conj(np.fft.fft(np.conj(P))) / M
    return c * p[:n]
