"""Drop-in for the reference's Nets.MLP (Nets.py:11-35): fc1 -> leaky_relu -> fc2 -> leaky_relu ->
fc3 -> leaky_relu, returning (fc3_map, fc2_map).  Same parameter names (fc1/fc2/fc3 nn.Linear), so a
reference state_dict loads unchanged; the forward pass is ONE fused CUDA kernel on the tcgen05
tensor cores (dm_mlp_forward_bf16: bf16 operands, fp32 accumulation in TMEM).  Inference only:
the kernel has no backward (the reference trains this toy network on MNIST, MLP.py, out of scope).

`dims` generalises the layer widths for the pair-MLP scorer of the merge path:
MLP(dims=(2 * D, 250, 2)).
"""
from __future__ import annotations

import torch
from torch import nn

from .raster import PackedMLP, mlp_forward


class MLP(nn.Module):
    def __init__(self, dims=(784, 250, 10)):
        super().__init__()
        n_in, hidden, n_out = dims
        self.fc1 = nn.Linear(n_in, hidden)          # Nets.py:14-16 (784 -> 250)
        self.fc2 = nn.Linear(hidden, hidden)        # Nets.py:18-20
        self.fc3 = nn.Linear(hidden, n_out)         # Nets.py:22-24 (250 -> 10)
        self._packed = None
        self._stamp = None

    def packed(self) -> PackedMLP:
        """bf16 tensor-core image of the current weights (re-packed when a parameter changes)."""
        ps = [self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias, self.fc3.weight, self.fc3.bias]
        stamp = tuple((p.data_ptr(), p._version) for p in ps)
        if self._packed is None or stamp != self._stamp:
            self._packed = PackedMLP(*ps)
            self._stamp = stamp
        return self._packed

    def forward(self, x):
        if not x.is_cuda:
            raise ValueError("deepmerge_b200.Nets.MLP needs CUDA tensors (there is no CPU path)")
        with torch.no_grad():
            fc3_map, fc2_map = mlp_forward(x.reshape(x.shape[0], -1), self.packed())
        return fc3_map, fc2_map
