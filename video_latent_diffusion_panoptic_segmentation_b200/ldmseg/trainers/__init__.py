from .trainers_ldm_cond import TrainerDiffusion

__all__ = ["TrainerDiffusion"]
