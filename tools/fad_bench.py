"""Single-GPU timing of dm_fad_moments (sum x x^T on tcgen05 + TMA) for ncu captures and roofline numbers."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diffmusic_b200 import fad  # noqa: E402

n, d = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (510976, 768)
engine = sys.argv[3] if len(sys.argv) > 3 else "tcgen05"
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
X = (torch.randn(n, d, device="cuda") * 0.6 + 0.25).half()
m = fad.EmbeddingMoments(d, engine=engine)
for _ in range(3):
    m.update(X)
torch.cuda.synchronize()
s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(iters):
    m.update(X)
t.record()
torch.cuda.synchronize()
ms = s.elapsed_time(t) / iters
peak = 1665.4
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["bf16_tflops"])
tf = 2.0 * n * d * d / (ms * 1e-3) / 1e12
print(json.dumps({"kernel": "fad_xtx_tc_kernel + fad_colsum_kernel", "engine": engine, "N": n, "d": d, "ms": ms,
                  "tflops": tf, "peak_bf16_tflops_measured": peak, "frac_of_dense_16bit_peak": tf / peak}))
