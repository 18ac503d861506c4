"""Helpers mirrored from phantom/utils/__init__.py that the hot-path API touches."""
from ..encoders import flatten  # noqa: F401
from . import samplers  # noqa: F401
