"""Fused host -> device staging of a step's sparse features: the Python mirror of
HbH2DTransferN (hybridbackend/tensorflow/ops/transfer/transfer.cc:68-80)."""
import torch

from hybridbackend_b200 import _lib
from hybridbackend_b200 import _util


def h2d_transfer_n(inputs, outputs=None, device=None):
  """inputs: list of host tensors (any dtype / shape, contiguous); returns the list of
  device tensors holding the same values.  Pinned inputs travel in ONE kernel launch
  that reads the host buffers over PCIe; pageable ones fall back to a copy each.
  Enqueued on the current stream; `outputs` (optional) are reused when given."""
  inputs = list(inputs)
  dev = torch.device(device if device is not None else f'cuda:{torch.cuda.current_device()}')
  if outputs is None:
    outputs = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in inputs]
  outputs = list(outputs)
  if len(outputs) != len(inputs):
    raise ValueError('h2d_transfer_n: inputs and outputs differ in length')
  for i, (h, d) in enumerate(zip(inputs, outputs)):
    if h.is_cuda or not h.is_contiguous():
      raise ValueError(f'h2d_transfer_n: input {i} must be a contiguous host tensor')
    _util.require_cuda(d, f'h2d_transfer_n: output {i}')
    if d.dtype != h.dtype or d.numel() != h.numel() or not d.is_contiguous():
      raise ValueError(f'h2d_transfer_n: output {i} does not match its input')
  n = len(inputs)
  if n == 0:
    return outputs
  with torch.cuda.device(dev):
    _lib.check(_lib.lib().hbH2DTransferN(
        n, _lib.ptr_array([t.data_ptr() for t in inputs]), _lib.ptr_array([t.data_ptr() for t in outputs]),
        _lib.i64_array([t.numel() * t.element_size() for t in inputs]), _util.stream_ptr()), 'h2d_transfer_n')
  return outputs
