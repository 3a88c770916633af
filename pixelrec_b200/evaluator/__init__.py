from .collector import Collector, DataStruct
from .evaluator import Evaluator
from .metrics import METRICS, metric_types, smaller_metrics

__all__ = ["Collector", "DataStruct", "Evaluator", "METRICS", "metric_types", "smaller_metrics"]
