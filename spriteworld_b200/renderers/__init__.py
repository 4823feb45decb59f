"""Renderers (reference: spriteworld/renderers/__init__.py)."""
from spriteworld_b200.renderers import color_maps  # noqa: F401
