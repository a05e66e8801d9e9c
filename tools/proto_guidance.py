"""NumPy (float64) prototype of the fused guidance math: T_mel fwd + loss + hand-derived VJP, phase variants, FIR
resample adjoint, overlap-save RIR correlation + adjoint.  Checked against torch autograd of the CPU oracle."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import operators as oo
from tests import stubs

NFFT, HOP, NB, NM = 1024, 160, 513, 64
hann = torch.hann_window(1024, periodic=True).double().numpy()
import torchaudio
fb = torchaudio.functional.melscale_fbanks(513, 0.0, 8000.0, 64, 16000, None, "htk").double().numpy()

def reflect_index(i, L):   # padded index i (0..L+1023) -> source index
    j = i - 512
    if j < 0: j = -j
    if j >= L: j = 2 * (L - 1) - j
    return j

def stft_guidance(y, ref, mode, window):
    """y: (L',) float64; ref: target in the compared space.  Returns loss, dL/dy.
    mode: 'mel_db' | 'mel_db_noclamp' | 'phase_mel' | 'phase_wav'"""
    L = len(y); T = 1 + L // HOP
    yp = np.array([y[reflect_index(i, L)] for i in range(L + 1024)])
    frames = np.stack([yp[t * HOP: t * HOP + NFFT] * window for t in range(T)])   # (T, 1024)
    X = np.fft.rfft(frames, axis=1)                                             # (T, 513)
    if mode.startswith("mel_db"):
        P = X.real ** 2 + X.imag ** 2
        mel = P @ fb                                                            # (T, 64)
        D = 10 * np.log10(np.maximum(mel, 1e-10))
        out = np.clip(D, -80, 80) if mode == "mel_db" else D
        d = ref.T - out
        loss = np.sqrt((d ** 2).sum())
        G = -d / loss
        if mode == "mel_db":
            G = G * ((D >= -80) & (D <= 80))
        melbar = G * (10 / np.log(10)) / np.maximum(mel, 1e-10) * (mel >= 1e-10)
        Pbar = melbar @ fb.T
        Xbar = 2 * Pbar * X
    elif mode == "phase_mel":
        mag = np.abs(X)
        mel = mag @ fb
        out = np.clip(mel, -80, 80)
        d = ref.T - out
        loss = np.sqrt((d ** 2).sum())
        G = -d / loss * ((mel >= -80) & (mel <= 80))
        magbar = G @ fb.T
        Xbar = magbar * X / np.where(mag > 0, mag, 1) * (mag > 0)
    else:
        mag = np.abs(X)
        d = ref.T - mag
        loss = np.sqrt((d ** 2).sum())
        magbar = -d / loss
        Xbar = magbar * X / np.where(mag > 0, mag, 1) * (mag > 0)
    # rfft adjoint (autograd convention): f[n] = Re sum_k Xbar_k e^{+2 pi i k n/N}
    Y = Xbar / 2; Y[:, 0] = Xbar[:, 0].real; Y[:, 512] = Xbar[:, 512].real
    fbar = np.fft.irfft(Y, n=1024, axis=1) * 1024 * window
    ypbar = np.zeros(L + 1024)
    for t in range(T):
        ypbar[t * HOP: t * HOP + NFFT] += fbar[t]
    ybar = np.zeros(L)
    for i in range(L + 1024):
        ybar[reflect_index(i, L)] += ypbar[i]
    return loss, ybar

def check(name, got, want, tol=1e-9):
    err = np.linalg.norm(got - want) / np.linalg.norm(want)
    print(f"{name}: rel {err:.2e}"); assert err < tol, name

L = 4000
w = stubs.synth_clips(1, L).double(); r = stubs.synth_clips(1, L, first=50).double()
# oracle modules in float64
import torchaudio.transforms as T
wav2mel = torch.nn.Sequential(T.MelSpectrogram(16000, 1024, 1024, 160, n_mels=64, power=2.0), T.AmplitudeToDB("power")).double()
mag2mel = T.MelScale(64, 16000, n_stft=513).double()

def torch_case(fwd, tr, x, meas, wav_space=False):
    x = x.clone().requires_grad_(True)
    p = fwd(x)
    diff = (meas - p) if wav_space else (tr(meas) - tr(p))
    loss = torch.linalg.norm(diff)
    return loss.item(), torch.autograd.grad(loss, x)[0][0].numpy()

ident = lambda x: x
tmel = lambda x: torch.clamp(wav2mel(x), -80, 80)
l, g = torch_case(ident, tmel, w, r)
l2, g2 = stft_guidance(w[0].numpy(), tmel(r)[0].numpy(), "mel_db", hann)
check("mel_db loss", np.array([l2]), np.array([l])); check("mel_db grad", g2, g)
l, g = torch_case(ident, wav2mel, w, r)
l2, g2 = stft_guidance(w[0].numpy(), wav2mel(r)[0].numpy(), "mel_db_noclamp", hann)
check("mel_db_noclamp grad", g2, g)
# quiet signal to exercise clamp / amin edges
wq = w * 3e-5
l, g = torch_case(ident, tmel, wq, r)
l2, g2 = stft_guidance(wq[0].numpy(), tmel(r)[0].numpy(), "mel_db", hann)
check("mel_db quiet grad", g2, g)
phase = lambda x: torch.abs(torch.stft(x, 1024, 160, 1024, return_complex=True))
pmel = lambda m: torch.clamp(mag2mel(m), -80, 80)
l, g = torch_case(phase, pmel, w, phase(r))
l2, g2 = stft_guidance(w[0].numpy(), pmel(phase(r))[0].numpy(), "phase_mel", np.ones(1024))
check("phase_mel loss", np.array([l2]), np.array([l])); check("phase_mel grad", g2, g)
l, g = torch_case(phase, None, w, phase(r), wav_space=True)
l2, g2 = stft_guidance(w[0].numpy(), phase(r)[0].numpy(), "phase_wav", np.ones(1024))
check("phase_wav grad", g2, g)

# ---------------- FIR resample fwd + gather-form adjoint
def resample_fwd(x, kern, orig, new, width):
    # kern: (new, W)
    L = len(x); W = kern.shape[1]
    nj = (L + width + width + orig - W) // orig + 1
    tgt = int(np.ceil(new * L / orig))
    y = np.zeros(nj * new)
    for j in range(nj):
        for p in range(new):
            acc = 0.0
            for k in range(W):
                i = orig * j + k - width
                if 0 <= i < L: acc += x[i] * kern[p, k]
            y[j * new + p] = acc
    return y[:tgt]

def resample_adj(ybar, kern, orig, new, width, L):
    W = kern.shape[1]; tgt = len(ybar)
    xbar = np.zeros(L)
    for i in range(L):
        # need j with 0 <= i + width - orig*j < W  ->  j in [ceil((i+width-W+1)/orig), floor((i+width)/orig)]
        jlo = max(0, -((-(i + width - W + 1)) // orig)); jhi = (i + width) // orig
        acc = 0.0
        for j in range(jlo, jhi + 1):
            k = i + width - orig * j
            for p in range(new):
                o = j * new + p
                if o < tgt: acc += ybar[o] * kern[p, k]
        xbar[i] = acc
    return xbar

for sr_new, Lr in ((8000, 1000), (1600, 1003), (12000, 801)):
    rs = T.Resample(16000, sr_new)  # float32 kernel
    import math
    gcd = math.gcd(16000, sr_new); orig, new = 16000 // gcd, sr_new // gcd
    kern = rs.kernel[:, 0, :].double().numpy(); width = rs.width
    x = torch.randn(1, Lr, dtype=torch.float64)
    xx = x.clone().requires_grad_(True)
    rsd = T.Resample(16000, sr_new, dtype=torch.float64); rsd.kernel = rs.kernel.double()
    y = rsd(xx)
    yb = torch.randn_like(y)
    (gx,) = torch.autograd.grad((y * yb).sum(), xx)
    y2 = resample_fwd(x[0].numpy(), kern, orig, new, width)
    check(f"resample fwd {sr_new}", y2, y[0].detach().numpy())
    check(f"resample adj {sr_new}", resample_adj(yb[0].numpy(), kern, orig, new, width, Lr), gx[0].numpy())

# ---------------- RIR correlation by overlap-save real FFT, + adjoint
def corr_os(x, h, NF):
    """y[i] = sum_k xz[i+k] h[k], xz zero-padded K//2 both sides, i = 0..L+2*(K//2)-K  (torch conv1d)."""
    L = len(x); K = len(h); pad = K // 2
    nout = L + 2 * pad - K + 1
    V = NF - K + 1                         # valid outputs per block
    Hc = np.conj(np.fft.rfft(h, NF))       # correlation = multiply by conj(H)
    y = np.zeros(nout)
    for b in range((nout + V - 1) // V):
        i0 = b * V                         # first output of this block; needs xz[i0 .. i0+NF-1] = x[i0-pad ...]
        seg = np.zeros(NF)
        for n in range(NF):
            s = i0 + n - pad
            if 0 <= s < L: seg[n] = x[s]
        out = np.fft.irfft(np.fft.rfft(seg) * Hc, NF)
        m = min(V, nout - i0)
        y[i0:i0 + m] = out[:m]
    return y

def corr_adj_os(ybar, h, NF, L):
    """xbar[j] = sum_k ybar[j + pad - k] h[k]  (true convolution with h, offset pad)."""
    K = len(h); pad = K // 2; nout = len(ybar)
    V = NF - K + 1
    Hf = np.fft.rfft(h, NF)
    xbar = np.zeros(L)
    for b in range((L + V - 1) // V):
        j0 = b * V
        # outputs j0..j0+V-1 need ybar[j + pad - k], k=0..K-1 -> ybar indices j0+pad-(K-1) .. j0+V-1+pad
        base = j0 + pad - (K - 1)
        seg = np.zeros(NF)
        for n in range(NF):
            s = base + n
            if 0 <= s < nout: seg[n] = ybar[s]
        out = np.fft.irfft(np.fft.rfft(seg) * Hf, NF)   # circular conv; valid at n >= K-1
        m = min(V, L - j0)
        xbar[j0:j0 + m] = out[K - 1:K - 1 + m]
    return xbar

for K, Lc in ((800, 5000), (801, 4097), (5000, 9000)):
    h = torch.randn(1, K, dtype=torch.float64); x = torch.randn(1, Lc, dtype=torch.float64)
    xx = x.clone().requires_grad_(True)
    y = torch.nn.functional.conv1d(xx.unsqueeze(1), h.unsqueeze(1), padding=K // 2).squeeze(1)
    yb = torch.randn_like(y)
    (gx,) = torch.autograd.grad((y * yb).sum(), xx)
    check(f"corr fwd K{K}", corr_os(x[0].numpy(), h[0].numpy(), 8192), y[0].detach().numpy())
    check(f"corr adj K{K}", corr_adj_os(yb[0].numpy(), h[0].numpy(), 8192, Lc), gx[0].numpy())
print("ALL OK")
