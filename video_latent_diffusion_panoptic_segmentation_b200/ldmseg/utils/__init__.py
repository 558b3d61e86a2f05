from .utils import OutputDict, get_rank, get_world_size, gpu_gather, is_dist_avail_and_initialized, is_main_process

__all__ = ["OutputDict", "get_rank", "get_world_size", "gpu_gather", "is_dist_avail_and_initialized",
           "is_main_process"]
