"""Minimal stand-in for the two matplotlib modules the reference's sprite.py uses."""
__version__ = '0-standin'
