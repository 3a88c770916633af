from .enum_type import EvaluatorType, InputType
from .utils import (barrier, calculate_valid_score, dict2str, early_stopping, ensure_dir, get_local_time, get_model,
                    get_rank, get_world_size, init_seed, set_color)

__all__ = ["InputType", "EvaluatorType", "get_model", "init_seed", "early_stopping", "calculate_valid_score", "dict2str",
           "ensure_dir", "get_local_time", "get_rank", "get_world_size", "barrier", "set_color"]
