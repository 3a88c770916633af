"""Enums of the plugin contract (reference: REC/utils/enum_type.py:3-17)."""
from enum import Enum


class InputType(Enum):
    SEQ = 1
    PAIR = 2
    AUGSEQ = 3


class EvaluatorType(Enum):
    RANKING = 1
    VALUE = 2
