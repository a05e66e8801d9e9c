"""NumPy emulation of the device algorithms (index math per thread), validated against np.fft / torch autograd
before transcribing to CUDA.  Development aid only."""
import numpy as np

R = 8

def stockham_fft(z, inverse=False):
    """Complex FFT, N = 8^P, radix-8 Stockham, thread j handles inputs j + r*T (T=N/8)."""
    N = len(z)
    T = N // R
    sign = +1.0 if inverse else -1.0
    src = z.astype(np.complex128).copy()
    Ns = 1
    while Ns < N:
        dst = np.zeros_like(src)
        for j in range(T):
            k = j % Ns
            v = np.array([src[j + r * T] for r in range(R)])
            # twiddle W_{Ns*R}^{r*k}
            for r in range(R):
                v[r] *= np.exp(sign * 2j * np.pi * r * k / (Ns * R))
            # 8-point DFT
            out = np.array([sum(v[r] * np.exp(sign * 2j * np.pi * r * q / R) for r in range(R)) for q in range(R)])
            j0 = (j // Ns) * Ns * R + k
            for q in range(R):
                dst[j0 + q * Ns] = out[q]
        src = dst
        Ns *= R
    return src

for N in (64, 512):
    z = np.random.randn(N) + 1j * np.random.randn(N)
    assert np.allclose(stockham_fft(z), np.fft.fft(z)), N
    assert np.allclose(stockham_fft(z, True), np.fft.ifft(z) * N), N
print("stockham ok")

def rfft_via_half(x):
    """1024-real FFT via 512-complex: returns X[0..N/2]."""
    N = len(x); H = N // 2
    z = x[0::2] + 1j * x[1::2]
    Z = stockham_fft(z)
    X = np.zeros(H + 1, dtype=np.complex128)
    for k in range(H + 1):
        Zk = Z[k % H]
        Zc = np.conj(Z[(H - k) % H])
        E = 0.5 * (Zk + Zc)
        O = -0.5j * (Zk - Zc)
        X[k] = E + np.exp(-2j * np.pi * k / N) * O
    return X

x = np.random.randn(512 * 2)
assert np.allclose(rfft_via_half(x), np.fft.rfft(x))
print("rfft ok")

def rfft_adjoint_via_half(Xb):
    """autograd adjoint of rfft: f[n] = Re sum_{k=0}^{H} Xb[k] e^{+2 pi i k n / N}; via 512-complex inverse FFT."""
    H = len(Xb) - 1; N = 2 * H
    # hermitian half-spectrum Y for unnormalised irfft: Y0 = Re Xb0, YH = Re XbH, Yk = Xb_k / 2
    Y = Xb.astype(np.complex128) / 2
    Y[0] = Xb[0].real; Y[H] = Xb[H].real
    Z = np.zeros(H, dtype=np.complex128)
    for k in range(H):
        Yk = Y[k]; Yc = np.conj(Y[H - k])
        Z[k] = (Yk + Yc) + 1j * np.exp(2j * np.pi * k / N) * (Yk - Yc)
    z = stockham_fft(Z, inverse=True)
    f = np.zeros(N)
    f[0::2] = z.real; f[1::2] = z.imag
    return f

Xb = np.random.randn(513) + 1j * np.random.randn(513)
n = np.arange(1024)
direct = np.array([np.real(np.sum(Xb * np.exp(2j * np.pi * np.arange(513) * nn / 1024))) for nn in n])
assert np.allclose(rfft_adjoint_via_half(Xb), direct)
print("rfft adjoint ok")
