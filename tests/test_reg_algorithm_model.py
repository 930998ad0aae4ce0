"""CPU restatement (numpy) of the arithmetic of stage_reg_kernel (pfft_b200/csrc/fft_reg_kernel.h), formula by formula:
the radix-3 decimation-in-frequency prologue with its thread / element mapping, the r2c post-processing by pairs and
the c2r pre-processing through forward butterflies on swapped parts -- each against numpy's FFT.  The GPU parity tests
check the kernel itself; this pins the algebra (and the index maps) where no GPU is available."""
import numpy as np
import pytest


def swap(z):
    return z.imag + 1j * z.real


@pytest.mark.parametrize("NSUB,E", [(64, 8), (128, 8), (256, 8), (512, 8)])
def test_radix3_prologue_and_output_distribution(NSUB, E):
    Q, NL, TS = 3, 3 * NSUB, NSUB // E
    TL = Q * TS
    rng = np.random.default_rng(NSUB)
    x = rng.standard_normal(NL) + 1j * rng.standard_normal(NL)
    w_nl = np.exp(-2j * np.pi * np.arange(NL) / NL)
    h = 0.86602540378443864676
    X = np.zeros(NL, dtype=complex)
    seen = np.zeros(NL, dtype=int)
    for r in range(TL):                       # thread r = Q t + q of the line
        t, q = divmod(r, Q)
        y = np.zeros(NSUB, dtype=complex)
        for j in range(NSUB):                 # (the thread only forms its own E points j = t + e TS; all of them here)
            a0, a1, a2 = x[j], x[j + NSUB], x[j + 2 * NSUB]
            s, d = a1 + a2, a1 - a2
            if q == 0:
                y[j] = a0 + s
            else:
                hh = h if q == 1 else -h
                y[j] = complex(a0.real - 0.5 * s.real + hh * d.imag, a0.imag - 0.5 * s.imag - hh * d.real) * w_nl[j * q]
        Y = np.fft.fft(y)                     # the NSUB-point Stockham passes
        for e in range(E):                    # element e of thread (t, q): sub-transform output k' = t + e TS ...
            kp = t + e * TS
            k = r + e * TL                    # ... is output r + e TL of the line (natural strided distribution)
            assert k == Q * kp + q
            X[k] = Y[kp]
            seen[k] += 1
    assert (seen == 1).all()
    np.testing.assert_allclose(X, np.fft.fft(x), rtol=0, atol=1e-10 * np.abs(x).sum())


@pytest.mark.parametrize("M", [64, 192, 512, 768])
def test_r2c_post_processing_by_pairs(M):
    n = 2 * M
    rng = np.random.default_rng(M)
    x = rng.standard_normal(n)
    Z = np.fft.fft(x[0::2] + 1j * x[1::2])                 # packed half-length transform
    w = np.exp(-2j * np.pi * np.arange(M + 1) / n)          # twh
    zl = np.concatenate([Z, Z[:1]])                         # Z[0] once more behind the line (mirror index M - 0)
    X = np.zeros(M + 1, dtype=complex)
    for k in range(M // 2):
        a, b = zl[k], zl[M - k]
        Ee = complex(0.5 * (a.real + b.real), 0.5 * (a.imag - b.imag))
        D = complex(0.5 * (a.real - b.real), 0.5 * (a.imag + b.imag))
        P = complex(D.imag * w[k].real + D.real * w[k].imag, D.imag * w[k].imag - D.real * w[k].real)   # (D / i) w^k
        X[k] = Ee + P
        X[M - k] = complex(Ee.real - P.real, P.imag - Ee.imag)                                           # conj(E - P)
    X[M // 2] = np.conj(Z[M // 2])
    np.testing.assert_allclose(X, np.fft.rfft(x), rtol=0, atol=1e-11 * n)


@pytest.mark.parametrize("M", [64, 192, 512, 768])
def test_c2r_pre_processing_and_swapped_butterflies(M):
    n = 2 * M
    rng = np.random.default_rng(M + 1)
    xr = rng.standard_normal(n)
    X = np.fft.rfft(xr)                                     # Hermitian half spectrum, M + 1 entries
    w = np.exp(-2j * np.pi * np.arange(M + 1) / n)
    S = X.copy()                                            # the staging line; pairs (k, M - k) are updated in place
    for k in range(M // 2 + 1):
        a, b = S[k], S[M - k]
        if k == 0:
            a, b = complex(a.real, 0.0), complex(b.real, 0.0)
        Ee = complex(a.real + b.real, a.imag - b.imag)
        D = complex(a.real - b.real, a.imag + b.imag)
        O = complex(D.real * w[k].real + D.imag * w[k].imag, D.imag * w[k].real - D.real * w[k].imag)   # D conj(w^k)
        S[k] = complex(Ee.imag + O.real, Ee.real - O.imag)                  # Zf[k], parts swapped
        if k != 0 and 2 * k != M:
            S[M - k] = complex(O.real - Ee.imag, Ee.real + O.imag)          # Zf[M - k], parts swapped
    z = np.fft.fft(S[:M])                                   # forward butterflies on the swapped parts
    out = np.empty(n)
    out[0::2], out[1::2] = z.imag, z.real                   # swapped back on the way out
    np.testing.assert_allclose(out, np.fft.irfft(X, n) * n, rtol=0, atol=1e-10 * n)
    np.testing.assert_allclose(out, n * xr, rtol=0, atol=1e-10 * n)
