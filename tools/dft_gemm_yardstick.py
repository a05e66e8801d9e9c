"""Yardstick for the design decision "shared-memory FFT instead of DFT-as-GEMM on the tensor cores" (DESIGN.md section 4).

Times the library GEMM (cuBLAS through torch.matmul) of exactly the contraction a DFT-GEMM STFT needs for BASELINE
config 2 -- frames (8016 x 1024) times the windowed DFT matrix (1024 x 1026: cos | -sin for 513 bins) -- in fp32,
TF32 and bf16, and the spectrum error of each against a float64 FFT.  A guidance pass needs this contraction twice
(forward and the VJP back to the frames), 3 passes each for a 3xTF32 split; the fused FFT kernel does forward + mel + loss
+ VJP + overlap-add in one launch (see profiles/README.md for its time).

    python tools/dft_gemm_yardstick.py > gpurun_out/dft_gemm_yardstick.json
"""
import json
import math
import statistics

import torch

dev = torch.device("cuda", 0)
T, N, F = 8016, 1024, 513
g = torch.Generator(device="cpu").manual_seed(0)
frames = (0.1 * torch.randn(T, N, generator=g)).to(dev)
n = torch.arange(N, dtype=torch.float64)
k = torch.arange(F, dtype=torch.float64)
win = torch.hann_window(N, periodic=True, dtype=torch.float64)
ang = 2 * math.pi * torch.outer(n, k) / N
dft = torch.cat([torch.cos(ang) * win[:, None], -torch.sin(ang) * win[:, None]], dim=1)  # (1024, 1026) float64
want = torch.fft.rfft(frames.double().cpu() * win, dim=1)
want = torch.cat([want.real, want.imag], dim=1)


def run(a, b, iters=20):
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        out = a @ b
        t.record()
        torch.cuda.synchronize()
        ms.append(s.elapsed_time(t))
    return statistics.median(ms), out


rows = []
flop = 2.0 * T * N * 2 * F
for name, dt, tf32 in (("fp32 (no tensor cores)", torch.float32, False), ("tf32", torch.float32, True),
                       ("bf16", torch.bfloat16, False)):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    ms, out = run(frames.to(dt), dft.to(dev).to(dt))
    err = float(torch.linalg.norm(out.double().cpu() - want) / torch.linalg.norm(want))
    rows.append({"gemm": name, "us_one_contraction": ms * 1e3, "tflops": flop / (ms * 1e-3) / 1e12,
                 "spectrum_rel_l2_vs_fp64_fft": err})
torch.backends.cuda.matmul.allow_tf32 = False
tf32_us = rows[1]["us_one_contraction"]
print(json.dumps({"what": "cuBLAS DFT-GEMM yardstick, frames (8016 x 1024) x DFT (1024 x 1026)", "rows": rows,
                  "guidance_pass_estimates_us": {
                      "plain TF32 (outside the 1e-4 parity bound), forward + VJP": 2 * tf32_us,
                      "3xTF32 split (parity), forward + VJP": 6 * tf32_us},
                  "note": "mel projection, loss, window, overlap-add and the extra HBM round trips of a GEMM "
                          "formulation (4.1 MB of spectrum per clip) are not included"}, indent=1))
