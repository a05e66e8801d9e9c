"""Generate tests/golden/istft.npz by running the REFERENCE's own `mel_spectrogram_to_waveform_with_phase`.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_istft_golden.py

The method lives in /root/reference/diffmusic/pipelines/pipeline_musicldm.py:263-301, a module that imports `diffusers`
(not installed).  The method itself only needs torch and torchaudio and never touches `self`, so its source is cut out
of the reference file with `ast` (read-only, unmodified) and executed here; the fixture stores the seeded inputs'
recipe (tests/stubs.py style: regenerated from seeds by the tests) and the reference OUTPUTS.
"""
import ast
import os
import sys

import numpy as np
import torch
import torchaudio

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
SRC = "/root/reference/diffmusic/pipelines/pipeline_musicldm.py"
NAME = "mel_spectrogram_to_waveform_with_phase"


def reference_function(src=SRC, name=NAME):
    tree = ast.parse(open(src).read())
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"torch": torch, "torchaudio": torchaudio}
            exec(compile(mod, src, "exec"), ns)
            return ns[name]
    raise RuntimeError(f"{name} not found in {src}")


from tests.stubs import ISTFT_CASES, SPECTROGRAM_CASES, istft_inputs  # noqa: E402  (seeded inputs, rebuilt by the tests)

if __name__ == "__main__":
    fn = reference_function()
    torch.set_num_threads(8)
    out = {}
    for name, case in ISTFT_CASES.items():
        mel, phase = istft_inputs(name)
        # a (1, 513, T) phase is squeezed by the reference and broadcast over the batch; (B, 513, T) is used as is
        wav = fn(None, mel, phase, original_waveform_length=case[4])
        out[name] = wav.numpy().astype(np.float32)
        print(name, tuple(wav.shape), float(wav.abs().max()))
    # the companion forward, diffmusic/utils.py:11-20 (that module imports soundfile etc., hence the same extraction):
    # magnitude and phase of two seeded 0.25 s clips (the second one ends in a ragged last hop)
    w2s = reference_function("/root/reference/diffmusic/utils.py", "waveform_to_spectrogram")
    from tests import stubs
    for name, length in SPECTROGRAM_CASES.items():
        mag, phase = w2s(stubs.synth_clips(2, length, first=70))
        out[name + "_mag"], out[name + "_phase"] = mag.numpy(), phase.numpy()
        print(name, tuple(mag.shape))
    np.savez_compressed(os.path.join(HERE, "istft.npz"), **out)
