"""Global device, QuantTorch/device.py:2."""
import torch
device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
