"""improver_b200.nbhood — same module surface as improver/nbhood/__init__.py."""
from collections.abc import Iterable
from typing import List, Optional, Tuple, Union


def _as_iterable(obj):
    """None -> [], scalars / strings -> [obj], iterables unchanged
    (improver/utilities/common_input_handle.py:12-18)."""
    if obj is None:
        return []
    is_seq = isinstance(obj, Iterable) and not isinstance(obj, (str, bytes))
    return obj if is_seq else [obj]


def radius_by_lead_time(radii, lead_times=None) -> Tuple[Union[float, List[float]], Optional[List[int]]]:
    """Normalise the (radii, lead_times) pair given on a command line
    (improver/nbhood/__init__.py:12-56).

    Without lead times exactly one radius is allowed and it is returned as a
    float; with lead times both lists must have the same length and are returned
    as floats / ints."""
    radii_list = _as_iterable(radii) if radii is not None else radii
    if lead_times is None:
        if len(radii_list) != 1:
            raise ValueError("Multiple radii have been supplied but no associated lead times.")
        return float(radii_list[0]), None
    lead_list = _as_iterable(lead_times)
    if len(radii_list) != len(lead_list):
        raise ValueError("If leadtimes are supplied, it must be a list of equal length to a list of radii.")
    return [float(r) for r in radii_list], [int(t) for t in lead_list]
