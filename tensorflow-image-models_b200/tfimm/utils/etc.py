"""Small shape helpers.  Behavioural contract: reference tfimm/utils/etc.py:7-26."""
from collections.abc import Iterable


def to_2tuple(x):
    """``x`` -> ``(x, x)``; iterables are truncated to their first two items."""
    if isinstance(x, Iterable):
        return tuple(x)[:2]
    return (x, x)


def make_divisible(value, divisor, min_value=None, round_limit=0.9):
    """Rounds ``value`` to the nearest multiple of ``divisor`` (at least ``min_value``), bumping
    one step up when rounding would lose more than ``1 - round_limit`` of the value.
    This fixes every EfficientNet channel count (reference: tfimm/utils/etc.py:14-26)."""
    floor = min_value or divisor
    rounded = (int(value + divisor / 2) // divisor) * divisor
    result = max(floor, rounded)
    if result < round_limit * value:
        result += divisor
    return result
