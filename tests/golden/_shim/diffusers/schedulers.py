from oracle.ddim_base import DDIMSchedulerBase as DDIMScheduler  # noqa: F401
