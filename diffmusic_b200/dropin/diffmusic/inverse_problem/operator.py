"""Drop-in for the reference's `diffmusic.inverse_problem.operator` (run.py imports the operator classes from here)."""
from diffmusic_b200.operators import (BaseOperator, IdentityOperator, MusicDereverberationOperator,  # noqa: F401
                                      MusicInpaintingOperator, PhaseRetrievalOperator, StyleGuidanceOperator,
                                      SuperResolutionOperator)
