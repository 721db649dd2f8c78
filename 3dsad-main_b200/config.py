"""Layer shapes of the hot path (SURVEY.md section 8 layer table, VoteNet-convention
defaults [LINEAGE]) and seeded random-init parameters shared by the GPU path, the oracle
parity tests and bench.py (no checkpoints are available offline)."""
from __future__ import annotations

import numpy as np

# name -> (npoint, radius, nsample)
LAYER_CFG = {
    "sa1": (2048, 0.2, 64),
    "sa2": (1024, 0.4, 32),
    "sa3": (512, 0.8, 16),
    "sa4": (256, 1.2, 16),
    "agg": (256, 0.3, 16),        # vote aggregation; radius is per-cluster (base 0.3 unused when adaptive)
    "alpha": 1.0, "r_min": 0.1, "r_max": 1.2,
    "size_scale": 1.0, "size_clip": 2.0,      # size head: size = size_scale * exp(clip(y, -size_clip, size_clip)) [DECISION]
}


def mlp_channels(input_feature_dim: int = 1, seed_feat_dim: int = 256, vote_factor: int = 1):
    """Channel lists per stage; SA stages include the +3 relative-xyz inputs (use_xyz=True)."""
    return {
        "sa1": [input_feature_dim + 3, 64, 64, 128],
        "sa2": [128 + 3, 128, 128, 256],
        "sa3": [256 + 3, 128, 128, 256],
        "sa4": [256 + 3, 128, 128, 256],
        "fp1": [256 + 256, 256, 256],
        "fp2": [256 + 256, 256, 256],
        "vote": [seed_feat_dim, seed_feat_dim, seed_feat_dim, (3 + seed_feat_dim) * vote_factor],
        "agg": [seed_feat_dim + 3, 128, 128, 128],
        "size": [seed_feat_dim, 128, 3],          # size head: vote features at the cluster centres -> box size (3)
    }


def make_params(seed: int = 0, input_feature_dim: int = 1, bias_std: float = 0.05):
    """dict stage -> [(W (Cout,Cin) f32, b (Cout,) f32), ...]; W ~ N(0, 1/fan_in), BN folded."""
    rng = np.random.default_rng(seed)
    params = {}
    for name, ch in mlp_channels(input_feature_dim).items():
        layers = []
        for cin, cout in zip(ch[:-1], ch[1:]):
            W = (rng.standard_normal((cout, cin)) / np.sqrt(cin)).astype(np.float32)
            b = (rng.standard_normal(cout) * bias_std).astype(np.float32)
            layers.append((W, b))
        params[name] = layers
    # voting offsets should be small displacements, not O(1) jumps
    Wv, bv = params["vote"][-1]
    Wv[:3] *= np.float32(0.25)
    bv[:3] *= np.float32(0.25)
    # size head: keep the log-sizes O(0.3) so that the radii spread over the clamp range instead of saturating it
    Ws, bs = params["size"][-1]
    Ws *= np.float32(0.3)
    return params
