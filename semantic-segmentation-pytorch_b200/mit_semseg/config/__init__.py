from .defaults import _C as cfg  # noqa: F401
