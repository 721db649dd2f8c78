"""Shared point-wise MLP (+ max-pool) host side -- SURVEY.md section 8(a) row a6.

INTERIM (round-1 bring-up): the contraction below still runs through torch.matmul in
bf16 (cuBLAS) on top of this repo's own grouping kernels; the hand-written tcgen05
kernel (csrc/mlp.cu) replaces `_chain` once it is parity-green.  Storage precision is
already the final one: bf16 inputs / weights / inter-layer activations, fp32 accumulate,
fp32 bias + ReLU, fp32 output (tolerance 2e-2 vs the fp32 oracle, BASELINE north_star)."""
from __future__ import annotations

from typing import List, Tuple

import torch

from . import ops


class PreparedLayer:
    __slots__ = ("W", "b", "W_bf16", "cin", "cout")

    def __init__(self, W: torch.Tensor, b: torch.Tensor):
        self.W = W.contiguous()
        self.b = b.contiguous()
        self.W_bf16 = W.to(torch.bfloat16).contiguous()
        self.cout, self.cin = W.shape


def prepare_layers(layers: List[Tuple[torch.Tensor, torch.Tensor]]) -> List[PreparedLayer]:
    return [PreparedLayer(W, b) for (W, b) in layers]


def _chain(rows: torch.Tensor, layers: List[PreparedLayer], last_relu: bool) -> torch.Tensor:
    """rows (R,Cin) -> (R,Cout) fp32; bf16 storage between layers, fp32 accumulate."""
    h = rows.to(torch.bfloat16)
    n = len(layers)
    for i, L in enumerate(layers):
        y = (h @ L.W_bf16.t()).float() + L.b
        if i < n - 1 or last_relu:
            y = torch.relu(y)
        h = y.to(torch.bfloat16) if i < n - 1 else y
    return h


def sa_group_mlp(xyz, new_xyz, features, idx, radius, layers: List[PreparedLayer],
                 use_xyz: bool = True, normalize_xyz: bool = True) -> torch.Tensor:
    """Group (relative, optionally radius-normalised xyz ++ features) -> MLP -> max over nsample.
    xyz (B,N,3), new_xyz (B,P,3), features (B,C,N) | None, idx (B,P,S) -> (B,Cout,P) fp32."""
    B, P, S = idx.shape
    g = ops.grouping_operation(xyz.transpose(1, 2).contiguous(), idx)          # (B,3,P,S)
    g = g - new_xyz.transpose(1, 2).unsqueeze(-1)
    if normalize_xyz:
        g = g / (radius[:, None, :, None] if torch.is_tensor(radius) else float(radius))
    if features is not None:
        gf = ops.grouping_operation(features.contiguous(), idx)
        g = torch.cat([g, gf], dim=1) if use_xyz else gf
    rows = g.permute(0, 2, 3, 1).reshape(B * P * S, g.shape[1])
    y = _chain(rows, layers, last_relu=True)
    return y.view(B, P, S, -1).max(dim=2)[0].transpose(1, 2).contiguous()


def pointwise_mlp(x: torch.Tensor, layers: List[PreparedLayer], last_relu: bool = True) -> torch.Tensor:
    """x (B,C,n) -> (B,Cout,n) fp32 (FP modules, voting)."""
    B, C, n = x.shape
    rows = x.transpose(1, 2).reshape(B * n, C)
    y = _chain(rows, layers, last_relu)
    return y.view(B, n, -1).transpose(1, 2).contiguous()
