from .modules import RotateQuantizedLinear

__all__ = ["RotateQuantizedLinear"]
