"""improver_b200.nbhood — mirror of improver/nbhood/__init__.py."""
from collections.abc import Iterable
from typing import List, Optional, Tuple, Union


def _as_iterable(obj):
    """improver/utilities/common_input_handle.py:12-18."""
    if obj is None:
        return []
    if not isinstance(obj, Iterable) or isinstance(obj, (str, bytes)):
        obj = [obj]
    return obj


def radius_by_lead_time(radii, lead_times=None) -> Tuple[Union[float, List[float]], Optional[List[int]]]:
    """Parse radii / lead times (improver/nbhood/__init__.py:12-56)."""
    if lead_times is not None:
        lead_times = _as_iterable(lead_times)
    if radii is not None:
        radii = _as_iterable(radii)
    if lead_times is None:
        if not len(radii) == 1:
            raise ValueError("Multiple radii have been supplied but no associated lead times.")
        radius_or_radii = float(radii[0])
    else:
        if not len(radii) == len(lead_times):
            raise ValueError("If leadtimes are supplied, it must be a list of equal length to a list of radii.")
        radius_or_radii = [float(x) for x in radii]
        lead_times = [int(x) for x in lead_times]
    return radius_or_radii, lead_times
