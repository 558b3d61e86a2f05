from .cityscapes_pap_eval import CityscapesPanopticEvaluator
from .new_eval import aggregate, eval_window, vpq_eval

__all__ = ["CityscapesPanopticEvaluator", "vpq_eval", "eval_window", "aggregate"]
