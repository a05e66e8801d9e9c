"""Launch the export chain (mel_spectrogram_to_waveform_with_phase) a few times at BASELINE size, for an ncu launch list:

    ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/istft_launches.csv \
        python tools/istft_probe.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import diffmusic_b200 as dm  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda", 0)
mel = torch.rand(B, 1, 1001, 64, device=dev) * 6.0 - 1.0
ph = (torch.rand(1, 513, 1001, device=dev) * 2.0 - 1.0) * 3.14159
for _ in range(4):
    y = dm.mel_spectrogram_to_waveform_with_phase(mel, ph, original_waveform_length=160000)
torch.cuda.synchronize()
print(float(y.abs().max()))
