"""Typing aliases the reference's tests/utils.py imports from torchsde.types (annotations only)."""
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple, Union  # noqa: F401

import torch
from torch import nn

Tensor = torch.Tensor
Tensors = Sequence[Tensor]
TensorOrTensors = Union[Tensor, Tensors]
Scalar = Union[float, Tensor]
Vector = Union[Sequence[float], Tensor]
Module = nn.Module
Modules = Sequence[Module]
ModuleOrModules = Union[Module, Modules]
