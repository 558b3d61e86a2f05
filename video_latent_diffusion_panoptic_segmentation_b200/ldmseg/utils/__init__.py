from .utils import (OutputDict, color_map, get_rank, get_world_size, gpu_gather, is_dist_avail_and_initialized,
                    is_main_process)

__all__ = ["OutputDict", "color_map", "get_rank", "get_world_size", "gpu_gather", "is_dist_avail_and_initialized",
           "is_main_process"]
