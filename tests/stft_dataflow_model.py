"""Numpy model of the exact dataflow of the n_fft = 2048 / 1024 CUDA kernels (csrc/stft.cu stft2048_kernel,
csrc/stft2048_v2.cuh, csrc/stft2048_v3.cuh stft_v3_kernel):
real 2048-point FFT of a windowed frame as one 1024-point complex FFT (32 lanes x 32 registers,
two in-register 32-point passes with a twiddle + transpose in between) followed by the
even/odd split done on (k, 1024-k) pairs held by partner lanes.  Used by
tests/test_stft_dataflow_model.py to validate the index maps and twiddle tables on CPU."""
import numpy as np

M = 1024  # complex FFT length = n_fft / 2


def tables():
    k1 = np.arange(32)[:, None]
    n2 = np.arange(32)[None, :]
    t1 = np.exp(-2j * np.pi * (k1 * n2) / M)          # T1[k1][n2] = W_1024^(n2*k1)
    t2 = np.exp(-2j * np.pi * np.arange(512) / (2 * M))  # T2[k] = W_2048^k, k < 512
    return t1, t2


def frame_fft_model(xw):
    """xw: windowed real frame [2048] -> onesided spectrum X[0..1024] via the kernel's dataflow."""
    t1, t2 = tables()
    z = xw[0::2] + 1j * xw[1::2]                      # z[n], n < 1024
    v = z.reshape(32, 32)                             # v[n1][n2] = z[32*n1 + n2]; lane = n2, slot = n1
    w32 = np.exp(-2j * np.pi * np.outer(np.arange(32), np.arange(32)) / 32)
    V1 = w32 @ v                                      # per lane: 32-pt FFT over n1 -> [k1][n2]
    V1 = V1 * t1                                      # twiddle
    u = V1.T                                          # exchange: lane k1 now holds u[n2][k1] over n2
    Z = (w32 @ u)                                     # [k2][k1]: lane k1, slot k2 holds Z[k1 + 32*k2]
    Zl = Z.T                                          # Zl[lane k1][slot k2]
    X = np.zeros(M + 1, dtype=np.complex128)
    for lane in range(32):
        for i in range(16):
            k = lane + 32 * i
            if lane == 0:
                if i == 0:
                    continue
                pl, ps = 0, 32 - i                     # partner (lane, slot): k' = 1024 - 32 i
            else:
                pl, ps = 32 - lane, 31 - i             # k' = (32-lane) + 32 (31-i) = 1024 - k
            a, b = Zl[lane][i], Zl[pl][ps]
            kp = pl + 32 * ps
            assert kp == M - k
            E2 = a + np.conj(b)
            O2 = (a - np.conj(b)) / 1j
            T = t2[k] * O2
            X[k] = 0.5 * (E2 + T)
            X[kp] = 0.5 * np.conj(E2 - T)
    z0 = Zl[0][0]
    X[0] = z0.real + z0.imag
    X[M] = z0.real - z0.imag
    X[512] = np.conj(Zl[0][16])
    return X


def frame_pair_fft_model_1024(xw2):
    """The n_fft = 1024 item of stft_v3_kernel<1024>: two windowed real frames [2][1024] -> X[2][513].
    Lane n2 runs one 16-point FFT per frame over n1 (z[32 n1 + n2]), twiddle W_512^(n2 k1), one 32 x 32 transpose whose
    source slots are (frame, k1), then lane (frame, k1) runs a 32-point FFT over n2: slot k2 = Z_frame[k1 + 16 k2]; the
    even/odd split pairs lane (frame, k1) with lane (frame, (16 - k1) % 16), slot 31 - i (k1 = 0: slot 32 - i) and uses
    W_1024^k = W_1024^k1 * W_64^i."""
    H = 512
    w16 = np.exp(-2j * np.pi * np.outer(np.arange(16), np.arange(16)) / 16)
    w32 = np.exp(-2j * np.pi * np.outer(np.arange(32), np.arange(32)) / 32)
    slots = np.zeros((32, 32), dtype=np.complex128)           # [slot = frame * 16 + k1][lane n2] before the transpose
    for fr in range(2):
        z = xw2[fr][0::2] + 1j * xw2[fr][1::2]                 # 512 complex points
        v = z.reshape(16, 32)                                  # v[n1][n2]
        V1 = w16 @ v                                           # [k1][n2]
        V1 = V1 * np.exp(-2j * np.pi * np.outer(np.arange(16), np.arange(32)) / H)
        slots[fr * 16:(fr + 1) * 16] = V1
    u = slots.T                                                # lane L = (fr, k1) holds u[n2][L]
    Zl = (w32 @ u).T                                           # Zl[lane][slot k2] = Z_fr[k1 + 16 k2]
    X = np.zeros((2, H + 1), dtype=np.complex128)
    for lane in range(32):
        fr, k1 = lane // 16, lane % 16
        cl = np.exp(-2j * np.pi * k1 / 1024)
        for i in range(16):
            k = k1 + 16 * i
            if k1 == 0:
                if i == 0:
                    continue
                pl, ps = lane, 32 - i
            else:
                pl, ps = fr * 16 + (16 - k1), 31 - i
            a, b = Zl[lane][i], Zl[pl][ps]
            kp = (pl % 16) + 16 * ps
            assert kp == H - k
            E2 = a + np.conj(b)
            O2 = (a - np.conj(b)) / 1j
            T = cl * np.exp(-2j * np.pi * i / 64) * O2
            X[fr][k] = 0.5 * (E2 + T)
            X[fr][kp] = 0.5 * np.conj(E2 - T)
        if k1 == 0:
            z0 = Zl[lane][0]
            X[fr][0] = z0.real + z0.imag
            X[fr][H] = z0.real - z0.imag
            X[fr][H // 2] = np.conj(Zl[lane][16])
    return X


def power_line_slot(k, n_fft):
    "slot of bin k inside a frame's power line in stft_v3_kernel: one pad slot per 32 bins keeps the walk's lanes conflict free"
    return k + (k >> 5)
