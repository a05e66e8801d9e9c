"""Drop-in for the reference's `diffmusic.metrics.lsd` (eval.py:5,124-129): same class, GPU implementation.
`diffmusic/metrics` is a namespace directory in the reference, so `fad.py` / `kl.py` keep resolving from it."""
from diffmusic_b200.metrics import LogSpectralDistance  # noqa: F401
