from . import SR_evaluator, dynamic_evaluate, feedforward_evaluate, pointwise_evaluate

__all__ = ["SR_evaluator", "feedforward_evaluate", "dynamic_evaluate", "pointwise_evaluate"]
