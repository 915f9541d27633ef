"""vLLM backend: quantisation plugin only (text-generation plumbing is out of scope).

`register()` is the `vllm.general_plugins` entry point, same contract as
/root/reference/paroquant/inference/backends/vllm/__init__.py:6-9 -- idempotent; importing the
two modules registers `torch.ops.rotation.rotate` and the "paroquant" quantisation config.
"""


def register() -> None:
    import paroquant_b200.kernels.cuda  # noqa: F401
    import paroquant_b200.inference.backends.vllm.plugin  # noqa: F401
