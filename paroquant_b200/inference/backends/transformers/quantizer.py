"""HF transformers loader glue: `from_pretrained` on a converted ParoQuant checkpoint.

Mirror of /root/reference/paroquant/inference/backends/transformers/quantizer.py:30-115: a quantization config
registered under ``"paroquant"`` and an ``HfQuantizer`` that, before the weights are loaded, swaps every ``nn.Linear``
whose module name has a ``.qweight`` key in the checkpoint for a ``RotateQuantizedLinear`` -- visual towers and other
unquantised layers are left alone.  Differences from the reference, both deliberate: the checkpoint scan is shared
with ``paroquant_b200.checkpoint_io``; bf16 is accepted next to fp16 (the fused kernels compute in either, the
reference's AutoAWQ GEMM is fp16 only, quantizer.py:78-82).

Importing this module registers the config and the quantizer with transformers; the package ``__init__`` imports it when
transformers is installed (``register()``), so a plain ``from_pretrained`` finds it.
"""
from __future__ import annotations

import logging
import os
from typing import TYPE_CHECKING

import torch
import torch.nn as nn
from transformers.quantizers.auto import register_quantization_config, register_quantizer
from transformers.quantizers.base import HfQuantizer
from transformers.utils.quantization_config import QuantizationConfigMixin

from ....checkpoint_io import find_quantized_modules
from .modules import RotateQuantizedLinear

if TYPE_CHECKING:
    from transformers import PreTrainedModel

logger = logging.getLogger(__name__)


def _local_dir(model_path: str) -> str:
    if os.path.isdir(model_path):
        return model_path
    from huggingface_hub import snapshot_download   # quantizer.py:32
    return snapshot_download(model_path)


@register_quantization_config("paroquant")
class ParoQuantConfig(QuantizationConfigMixin):
    """`quantization_config` of a ParoQuant checkpoint (quantizer.py:47-66; written by convert.py:450-455)."""

    def __init__(self, bits: int = 4, group_size: int = 128, krot: int = 8, modules_to_not_convert: list[str] | None = None, **kwargs):
        self.quant_method = "paroquant"
        self.bits = bits
        self.group_size = group_size
        self.krot = krot
        self.modules_to_not_convert = modules_to_not_convert
        if hasattr(self, "post_init"):
            self.post_init()


def replace_quantized_linears(model: nn.Module, quantized_modules: set[str], *, bits: int = 4, group_size: int = 128, krot: int = 8) -> int:
    """Swap the named ``nn.Linear`` modules for ``RotateQuantizedLinear`` (quantizer.py:88-115); returns the count."""
    n = 0
    for name, module in list(model.named_modules()):
        if not isinstance(module, nn.Linear) or name not in quantized_modules:
            continue
        parent_name, _, attr = name.rpartition(".")
        parent = model.get_submodule(parent_name) if parent_name else model
        setattr(parent, attr, RotateQuantizedLinear(module.in_features, module.out_features, bias=module.bias is not None,
                                                    group_size=group_size, bits=bits, krot=krot))
        n += 1
    return n


@register_quantizer("paroquant")
class ParoQuantHfQuantizer(HfQuantizer):
    """Replaces nn.Linear with RotateQuantizedLinear for the layers the checkpoint quantises."""

    requires_calibration = True

    def validate_environment(self, **kwargs):
        if not torch.cuda.is_available():
            raise RuntimeError("ParoQuant requires CUDA.")              # quantizer.py:75-76

    def update_dtype(self, dtype):
        if dtype not in (torch.float16, torch.bfloat16):
            logger.warning("ParoQuant kernels compute in float16 or bfloat16. Overriding dtype=%s -> float16.", dtype)
            return torch.float16
        return dtype

    def _process_model_before_weight_loading(self, model: "PreTrainedModel", **kwargs):
        qcfg = self.quantization_config
        quantized = find_quantized_modules(_local_dir(model.config._name_or_path))
        if qcfg.modules_to_not_convert:
            quantized -= set(qcfg.modules_to_not_convert)
        n = replace_quantized_linears(model, quantized, bits=qcfg.bits, group_size=qcfg.group_size, krot=qcfg.krot)
        logger.info("Found %d quantized modules in checkpoint, replaced %d.", len(quantized), n)

    def _process_model_after_weight_loading(self, model: "PreTrainedModel", **kwargs):
        return model

    @property
    def is_trainable(self) -> bool:
        return False

    def is_serializable(self, **kwargs) -> bool:
        return True
