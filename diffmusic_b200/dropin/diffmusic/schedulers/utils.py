from diffmusic_b200.schedulers import InverseProblemSchedulerOutput  # noqa: F401
