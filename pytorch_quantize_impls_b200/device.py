"""Default device of the package (the reference keeps a module-level `device`, QuantTorch/device.py:2).

The quantized kernels exist for CUDA only, so the default is the current CUDA device when one is visible; on a machine
without a GPU the attribute is the CPU device and every quantized op raises on first use (there is no CPU fallback)."""
import torch


def default_device():
    """cuda:<current> when CUDA is available, else cpu."""
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


device = default_device()
