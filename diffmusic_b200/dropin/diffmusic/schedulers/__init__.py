"""Drop-in for the reference's `diffmusic.schedulers` (same import path, same names).

`diffmusic` is a PEP-420 namespace package in the reference (no diffmusic/__init__.py), so putting
diffmusic_b200/dropin ahead of the reference on sys.path swaps in these two sub-packages while run.py,
diffmusic/pipelines, diffmusic/constants.py ... keep resolving from the reference."""
from diffmusic_b200.schedulers import (DDIMScheduler, DiffMusicScheduler, DPSScheduler, DSGScheduler,  # noqa: F401
                                       InverseProblemSchedulerOutput, MPGDScheduler, get_scheduler)
