"""Tensor carriers handed across the public API.

The reference returns ``tf.Tensor``s; callers use ``.numpy()``, ``.shape`` and arithmetic on them.  Here large results are
``DeviceTensor`` - a ``torch.Tensor`` subclass living on the GPU whose ``.numpy()`` copies to the host, and which exports
``__dlpack__`` (inherited) so a TensorFlow/GPflow caller can take it without leaving the device - and small host-side results
(hyper-parameter matrices, scalars) are ``HostTensor``, an ``np.ndarray`` subclass with a ``.numpy()`` method.
"""
from __future__ import annotations

import numpy as np
import torch


class DeviceTensor(torch.Tensor):
    @staticmethod
    def wrap(t: torch.Tensor) -> 'DeviceTensor':
        return t.as_subclass(DeviceTensor)

    def numpy(self, *args, **kwargs) -> np.ndarray:   # tf.Tensor.numpy() semantics: always a host array
        return torch.Tensor.numpy(self.detach().as_subclass(torch.Tensor).cpu())

    def to_dlpack(self):
        return torch.utils.dlpack.to_dlpack(self.as_subclass(torch.Tensor))


class HostTensor(np.ndarray):
    def __new__(cls, value):
        return np.asarray(value, dtype=np.float64).view(cls)

    def numpy(self) -> np.ndarray:
        return np.asarray(self)


def as_device(a, device=None) -> torch.Tensor:
    """Anything array-like -> plain contiguous CUDA float64 torch.Tensor. Raises without a GPU: there is no CPU path."""
    if not torch.cuda.is_available():
        raise RuntimeError('romcomma B200 path needs a CUDA device: there is no CPU fallback.')
    if isinstance(a, torch.Tensor):
        return a.as_subclass(torch.Tensor).to(device=device or 'cuda', dtype=torch.float64).contiguous()
    if hasattr(a, 'numpy') and not isinstance(a, np.ndarray):
        a = a.numpy()
    return torch.as_tensor(np.require(a, dtype=np.float64, requirements=['C', 'W'])).to(device or 'cuda')
