"""Comparison operators (improver/utilities/probability_manipulation.py:15-66)."""
import operator
from collections import namedtuple
from typing import Dict


def comparison_operator_dict() -> Dict[str, namedtuple]:
    """String comparison operator -> (function, spp_string, inverse)."""
    comparison = namedtuple("inequality", "function, spp_string, inverse")
    table = {}
    for keys, fn, spp, inv in (
        (("ge", "GE", ">="), operator.ge, "greater_than_or_equal_to", "lt"),
        (("gt", "GT", ">"), operator.gt, "greater_than", "le"),
        (("le", "LE", "<="), operator.le, "less_than_or_equal_to", "gt"),
        (("lt", "LT", "<"), operator.lt, "less_than", "ge"),
        (("eq", "EQ", "=="), operator.eq, "equal_to", "ne"),
    ):
        for k in keys:
            table[k] = comparison(function=fn, spp_string=spp, inverse=inv)
    return table
