"""Restated helpers of pytorchvideo.layers.utils (imported at /root/reference/model/x3d.py:16)."""
import math


def set_attributes(obj, params=None):
    # every constructor local except `self` becomes an attribute (registers sub-modules)
    for name, value in (params or {}).items():
        if name != "self":
            setattr(obj, name, value)


def round_width(width, multiplier, min_width=8, divisor=8, ceil=False):
    if not multiplier:
        return width
    scaled = width * multiplier
    floor_to = min_width or divisor
    if ceil:
        rounded = max(floor_to, int(math.ceil(scaled / divisor)) * divisor)
    else:
        rounded = max(floor_to, int(scaled + divisor / 2) // divisor * divisor)
    if rounded < 0.9 * scaled:
        rounded += divisor
    return int(rounded)


def round_repeats(repeats, multiplier):
    return repeats if not multiplier else int(math.ceil(multiplier * repeats))
