"""HF transformers backend.  Importing the package registers the "paroquant" quantizer with transformers when it is
installed (the reference does the same from its package __init__: paroquant/inference/backends/transformers/quantizer.py:30-44),
so `AutoModelForCausalLM.from_pretrained(<PARO checkpoint>)` works after `import paroquant_b200.inference.backends.transformers`.
"""
from .modules import RotateQuantizedLinear

__all__ = ["RotateQuantizedLinear", "register"]


def register() -> bool:
    """Register ParoQuantConfig / ParoQuantHfQuantizer with transformers; False when transformers is missing."""
    try:
        from . import quantizer  # noqa: F401  (decorators register on import)
    except ImportError:
        return False
    return True


register()
