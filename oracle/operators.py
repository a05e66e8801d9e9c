"""CPU restatement of diffmusic/inverse_problem/{operator,noise}.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Same library calls as the reference (torch.stft through torchaudio MelSpectrogram, MelScale, AmplitudeToDB,
Resample, F.conv1d), on CPU tensors, fp32.  Pinned against tests/golden/operators.npz, which was produced by the
reference's own classes (tests/golden/make_golden.py).
"""
from __future__ import annotations

import torch
import torchaudio.transforms as T


# ---------------------------------------------------------------------------------------------- transforms
_WAV2MEL = {}


def _wav2mel(sample_rate=16000):
    """operator.py:24-33 -- MelSpectrogram(n_fft 1024, hop 160, win 1024, 64 mels, power 2) + AmplitudeToDB(power)."""
    if sample_rate not in _WAV2MEL:
        _WAV2MEL[sample_rate] = torch.nn.Sequential(
            T.MelSpectrogram(sample_rate=sample_rate, n_fft=1024, hop_length=160, win_length=1024, n_mels=64,
                             power=2.0),
            T.AmplitudeToDB(stype="power"))
    return _WAV2MEL[sample_rate]


def mel_db(wav, clamp=True, sample_rate=16000):
    """T_mel.  operator.py:35-36 (clamped) / operator.py:123-124 (inpainting: NOT clamped)."""
    out = _wav2mel(sample_rate)(wav)
    return torch.clamp(out, min=-80, max=80) if clamp else out


_MAG2MEL = []


def phase_mel(mag):
    """PhaseRetrievalOperator.transform, operator.py:153-154: clamp(MelScale(64, 16000, 513)(mag.float()), +-80)."""
    if not _MAG2MEL:
        _MAG2MEL.append(T.MelScale(n_mels=64, sample_rate=16000, n_stft=1024 // 2 + 1))
    return torch.clamp(_MAG2MEL[0](mag.float()), min=-80, max=80)


# ---------------------------------------------------------------------------------------------- noise
def gaussian_noise(x, sigma):
    """noise.py:13-18 (draws from the global generator of x's device even when sigma == 0)."""
    return x + torch.randn_like(x) * sigma


# ---------------------------------------------------------------------------------------------- operators A(x)
def inpaint_mask(audio_length_in_s, sample_rate, mask_type, start_inpainting_s=None, end_inpainting_s=None,
                 mask_percentage=0.3, mask_duration_s=0.1, interval_s=1):
    """MusicInpaintingOperator.generate_mask, operator.py:87-121.  Integer index math; `random` consumes the global
    CPU generator through torch.randint exactly as the reference does."""
    n = audio_length_in_s * sample_rate
    mask = torch.ones([1, n])
    if mask_type == "box":
        if start_inpainting_s is not None and end_inpainting_s is not None:
            mask[:, int(start_inpainting_s * sample_rate): int(end_inpainting_s * sample_rate)] = 0.
    elif mask_type == "random":
        dur = int(mask_duration_s * sample_rate)
        count = max(1, int(mask_percentage * n) // dur)
        for _ in range(count):
            s = torch.randint(0, mask.shape[1] - dur, (1,))
            mask[:, s:s + dur] = 0.
    elif mask_type == "periodic":
        step = int(interval_s * sample_rate)
        dur = int(mask_duration_s * sample_rate)
        for s in range(0, mask.shape[1], step):
            mask[:, s:min(s + dur, mask.shape[1])] = 0.
    return mask


def a_inpaint(wav, mask, sigma=0.0):
    """operator.py:132-133."""
    return gaussian_noise(wav * mask.to(wav.device), sigma)


def a_phase(wav, n_fft=1024, hop_length=160, win_length=1024, sigma=0.0):
    """operator.py:162-171: |stft| with window=None (rectangular), center=True/reflect (torch.stft defaults)."""
    spec = torch.stft(wav, n_fft=n_fft, hop_length=hop_length, win_length=win_length, return_complex=True)
    return gaussian_noise(torch.abs(spec), sigma)


_RESAMPLERS = {}


def a_superres(wav, sample_rate=16000, scale=2, sigma=0.0):
    """operator.py:180,203-205: torchaudio Resample(sample_rate -> sample_rate // scale) on data.float()."""
    key = (sample_rate, scale)
    if key not in _RESAMPLERS:
        _RESAMPLERS[key] = T.Resample(orig_freq=sample_rate, new_freq=sample_rate // scale)
    return gaussian_noise(_RESAMPLERS[key](wav.float()), sigma)


def draw_impulse_response(ir_length=800, decay_factor=0.85):
    """operator.py:238-242: global-CPU-generator randn -> cumsum * decay -> / max|.| ; shape (1, K)."""
    ir = torch.randn(ir_length)
    ir = torch.cumsum(ir, dim=0) * decay_factor
    ir /= ir.abs().max()
    return ir.unsqueeze(0)


def a_dereverb(wav, ir, sigma=0.0):
    """operator.py:247-250 with the impulse response given (the reference redraws it on every forward call)."""
    out = torch.nn.functional.conv1d(wav.unsqueeze(1).float(), ir.to(wav.device).unsqueeze(1),
                                     padding=ir.size(1) // 2).squeeze(1)
    return gaussian_noise(out, sigma)


# ---------------------------------------------------------------------------------------------- operator objects
class OracleOperator:
    """forward / transform / inverse_transform triple with the reference's per-task wiring (run.py:159-212).

    kind: identity | inpainting | phase_retrieval | super_resolution | dereverberation
    For dereverberation `forward` draws a fresh impulse response per call (reference quirk, operator.py:246) unless
    `fixed_ir` is set; the last one used is kept in `self.last_ir`.
    """

    def __init__(self, kind, sigma=0.0, mask=None, scale=2, sample_rate=16000, ir_length=800, decay_factor=0.85,
                 n_fft=1024, hop_length=160, win_length=1024, fixed_ir=None):
        self.kind, self.sigma, self.mask, self.scale, self.sample_rate = kind, sigma, mask, scale, sample_rate
        self.ir_length, self.decay_factor, self.fixed_ir, self.last_ir = ir_length, decay_factor, fixed_ir, None
        self.n_fft, self.hop_length, self.win_length = n_fft, hop_length, win_length

    def forward(self, data, **kw):
        k = self.kind
        if k == "identity":
            return data
        if k == "inpainting":
            return a_inpaint(data, self.mask, self.sigma)
        if k == "phase_retrieval":
            return a_phase(data, self.n_fft, self.hop_length, self.win_length, self.sigma)
        if k == "super_resolution":
            return a_superres(data, self.sample_rate, self.scale, self.sigma)
        if k == "dereverberation":
            ir = self.fixed_ir if self.fixed_ir is not None else draw_impulse_response(self.ir_length,
                                                                                       self.decay_factor)
            self.last_ir = ir
            return a_dereverb(data, ir, self.sigma)
        raise ValueError(k)

    def transform(self, x):
        if self.kind == "phase_retrieval":
            return phase_mel(x)
        return mel_db(x, clamp=(self.kind != "inpainting"))

    def inverse_transform(self, mel, vocoder):
        """operator.py:38-42."""
        if mel.dim() == 4:
            mel = mel.squeeze(1)
        return vocoder(mel)
