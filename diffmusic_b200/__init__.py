"""diffmusic_b200 -- B200-native (sm_100a) implementation of DiffMusic's guidance hot path.

Host-side mirror of the reference's `diffmusic.inverse_problem` (operators, noise) and `diffmusic.schedulers`
(DDIM / DPS / MPGD / DSG / DiffMusic) plus the FAD embedding statistics, all running on hand-written CUDA kernels behind
the C ABI of include/dm_abi.h.  Importing the package does not need a GPU; calling any kernel does.
"""
from .noise import GaussianNoise, PoissonNoise, get_noiser  # noqa: F401
from .operators import (BaseOperator, IdentityOperator, MusicDereverberationOperator,  # noqa: F401
                        MusicInpaintingOperator, PhaseRetrievalOperator, StyleGuidanceOperator,
                        SuperResolutionOperator)
from .schedulers import (DDIMScheduler, DiffMusicScheduler, DPSScheduler, DSGScheduler,  # noqa: F401
                         InverseProblemSchedulerOutput, MPGDScheduler, get_scheduler)

from .graph import GraphedGuidedStep, HostPipelinedStep  # noqa: F401,E402
from .istft import mel_spectrogram_to_waveform_with_phase, waveform_to_spectrogram  # noqa: F401,E402
from .driver import BatchedGuidedSampler, BatchedSamplerOutput  # noqa: F401,E402

__version__ = "0.1.0"
