"""The four names the reference's tests/utils.py takes from `torchsde.types` (annotations only)."""
import typing

import torch

Callable = typing.Callable
Optional = typing.Optional
TensorOrTensors = typing.Union[torch.Tensor, typing.Sequence[torch.Tensor]]
ModuleOrModules = typing.Union[torch.nn.Module, typing.Sequence[torch.nn.Module]]
