from .ddim_scheduler import DDIMNoiseScheduler, DDIMNoiseSchedulerOutput

__all__ = ["DDIMNoiseScheduler", "DDIMNoiseSchedulerOutput"]
