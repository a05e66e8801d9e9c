"""Drop-in for the reference's `diffmusic.metrics.mse` (eval.py:6,130): same class, GPU implementation."""
from diffmusic_b200.metrics import MeanSquaredError  # noqa: F401
