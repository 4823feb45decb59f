"""Shape table and ShapeType ids (reference: spriteworld/constants.py:27-56)."""
import enum

import numpy as np

from spriteworld_b200 import shapes

# (builder, kwargs) per shape name, in ShapeType order
_SPEC = (
    ('triangle', shapes.polygon, dict(num_sides=3, theta_0=np.pi / 2)),
    ('square', shapes.polygon, dict(num_sides=4, theta_0=np.pi / 4)),
    ('pentagon', shapes.polygon, dict(num_sides=5, theta_0=np.pi / 2)),
    ('hexagon', shapes.polygon, dict(num_sides=6)),
    ('octagon', shapes.polygon, dict(num_sides=8)),
    ('circle', shapes.polygon, dict(num_sides=30)),
    ('star_4', shapes.star, dict(num_sides=4, theta_0=np.pi / 4)),
    ('star_5', shapes.star, dict(num_sides=5, theta_0=np.pi + np.pi / 10)),
    ('star_6', shapes.star, dict(num_sides=6)),
    ('spoke_4', shapes.spokes, dict(num_sides=4, theta_0=np.pi / 4)),
    ('spoke_5', shapes.spokes, dict(num_sides=5, theta_0=np.pi + np.pi / 10)),
    ('spoke_6', shapes.spokes, dict(num_sides=6)),
)

SHAPES = {name: fn(**kw) for name, fn, kw in _SPEC}

ShapeType = enum.IntEnum('ShapeType', [name for name, _, _ in _SPEC], start=1)
ShapeType.__doc__ = 'Integer ids of SHAPES, usable as a state description.'

SHAPE_NAMES = tuple(name for name, _, _ in _SPEC)
