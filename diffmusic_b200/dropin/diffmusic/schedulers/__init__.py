"""Drop-in for the reference's `diffmusic.schedulers` (same import path, same names).

`diffmusic` is a PEP-420 namespace package in the reference (no diffmusic/__init__.py), so putting
diffmusic_b200/dropin ahead of the reference on sys.path swaps in these two sub-packages while run.py,
diffmusic/pipelines, diffmusic/constants.py ... keep resolving from the reference."""
from pkgutil import extend_path

# Out-of-scope members of the reference package (scheduling_ditto.py: DITTO needs the UNet's backward) must stay
# importable under their reference paths (diffmusic/schedulers/__init__.py:19-20): this regular package would shadow the
# reference's directory, so that directory is appended to the search path -- modules defined here win, anything else
# (`diffmusic.schedulers.scheduling_ditto`) still resolves to the reference's file.
__path__ = extend_path(__path__, __name__)

from diffmusic_b200.schedulers import (DDIMScheduler, DiffMusicScheduler, DPSScheduler, DSGScheduler,  # noqa: F401
                                       InverseProblemSchedulerOutput, MPGDScheduler, get_scheduler)
