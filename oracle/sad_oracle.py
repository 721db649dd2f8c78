"""NumPy / torch-CPU restatement of the 3DSAD set-abstraction hot path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED: the
mounted reference is ``/root/reference/README.md:1-2`` and nothing else, so
every function below follows SURVEY.md section 8(a) (rows a1..a9) and the
arithmetic contract of section 7 H1/H2 instead of a reference file:line.

Arithmetic contract (H1/H2), shared with the CUDA kernels:
  * all distance math is IEEE fp32, one rounding per operation, NO fused
    multiply-add:  d2 = ((dx*dx) + (dy*dy)) + (dz*dz),  dx = p.x - q.x  (p is the
    candidate / scanned point, q the query / last selected point);
  * radius test is strict:  d2 < r*r  with r*r rounded once in fp32;
  * every argmax / argmin tie is broken towards the LOWEST index;
  * inputs are required to be finite.
NumPy evaluates each float32 ufunc with a single rounding and never contracts,
which is what makes this file the spec.
"""
from __future__ import annotations

import sys

import numpy as np

__all__ = [
    "sqdist", "furthest_point_sample", "gather_operation", "gather_operation_grad",
    "ball_query", "ball_query_adaptive", "grouping_operation", "grouping_operation_grad",
    "three_nn", "three_interpolate", "three_interpolate_grad", "interpolation_weights",
    "shared_mlp", "size_to_radius", "query_and_group", "sa_module", "fp_module",
    "voting_module", "vote_aggregation", "backbone_forward", "detector_hot_path",
    "bf16_round",
]

F32 = np.float32
FPS_INIT = F32(1e10)          # mind[k] = 1e10


def _f32c(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def sqdist(p, q):
    """d2 between broadcastable (...,3) fp32 arrays, contract order (H1)."""
    dx = p[..., 0] - q[..., 0]
    dy = p[..., 1] - q[..., 1]
    dz = p[..., 2] - q[..., 2]
    return ((dx * dx) + (dy * dy)) + (dz * dz)


# --------------------------------------------------------------------------- a1
def furthest_point_sample(xyz, npoint):
    """SURVEY 8(a) a1.  xyz (B,N,3) f32 -> idx (B,npoint) i32.

    sel[0] = 0; mind[:] = 1e10; each step mind = min(mind, d2(., last)), next =
    argmax(mind) with ties to the lowest index (np.argmax returns the first max).
    The lineage CUDA quirk of skipping near-origin points is deliberately NOT
    replicated (SURVEY a1 [DECISION])."""
    xyz = _f32c(xyz)
    B, N, _ = xyz.shape
    if not (1 <= npoint):
        raise ValueError("npoint must be >= 1")
    if N < 1:
        raise ValueError("N must be >= 1")
    out = np.zeros((B, npoint), dtype=np.int32)
    for b in range(B):
        pts = xyz[b]
        mind = np.full((N,), FPS_INIT, dtype=np.float32)
        last = 0
        for j in range(1, npoint):
            d = sqdist(pts, pts[last][None, :])
            np.minimum(mind, d, out=mind)
            last = int(np.argmax(mind))
            out[b, j] = last
    return out


# --------------------------------------------------------------------------- a2
def gather_operation(features, idx):
    """a2: out[b,c,j] = features[b,c,idx[b,j]].  (B,C,N),(B,npoint)->(B,C,npoint)."""
    features = np.asarray(features)
    idx = np.asarray(idx)
    B = features.shape[0]
    return np.stack([features[b][:, idx[b]] for b in range(B)], axis=0)


def gather_operation_grad(grad_out, idx, N):
    """a2 backward: scatter-add of grad_out (B,C,npoint) into (B,C,N)."""
    grad_out = np.asarray(grad_out)
    B, C, _ = grad_out.shape
    g = np.zeros((B, C, N), dtype=grad_out.dtype)
    for b in range(B):
        np.add.at(g[b], (slice(None), idx[b]), grad_out[b])
    return g


# ---------------------------------------------------------------------- a3 / a4
def _ball_query_impl(r2, nsample, xyz, new_xyz, chunk=128):
    xyz = _f32c(xyz)
    new_xyz = _f32c(new_xyz)
    B, N, _ = xyz.shape
    npoint = new_xyz.shape[1]
    out = np.zeros((B, npoint, nsample), dtype=np.int32)
    for b in range(B):
        pts = xyz[b]
        for q0 in range(0, npoint, chunk):
            q = new_xyz[b, q0:q0 + chunk]
            d2 = sqdist(pts[None, :, :], q[:, None, :])             # (Q,N)
            mask = d2 < r2[b, q0:q0 + chunk, None]
            for i in range(q.shape[0]):
                hits = np.flatnonzero(mask[i])
                if hits.size == 0:
                    continue                                        # all zeros
                k = min(nsample, hits.size)
                out[b, q0 + i, :] = hits[0]                         # first-hit padding
                out[b, q0 + i, :k] = hits[:k]
    return out


def ball_query(radius, nsample, xyz, new_xyz):
    """a3.  Scan candidates k ascending; hit iff d2 < r*r (strict); the first hit
    fills every slot, later hits fill slots in order, stop at nsample; no hit ->
    all zeros.  -> idx (B,npoint,nsample) i32."""
    B, npoint = np.asarray(new_xyz).shape[:2]
    r = F32(radius)
    r2 = np.full((B, npoint), r * r, dtype=np.float32)
    return _ball_query_impl(r2, nsample, xyz, new_xyz)


def ball_query_adaptive(radius_t, nsample, xyz, new_xyz):
    """a4 (3DSAD-specific).  As a3 but r = radius_t[b,j] per query / cluster."""
    radius_t = _f32c(radius_t)
    r2 = radius_t * radius_t
    return _ball_query_impl(r2, nsample, xyz, new_xyz)


def size_to_radius(size, alpha=1.0, r_min=0.1, r_max=1.2):
    """a4 helper [DECISION, unpinned]: r = clamp(alpha * 0.5 * ||size||_2, r_min, r_max).
    size (B,K,3) f32 -> (B,K) f32.  Evaluated as sqrt(((sx*sx)+(sy*sy))+(sz*sz))."""
    s = _f32c(size)
    n2 = ((s[..., 0] * s[..., 0]) + (s[..., 1] * s[..., 1])) + (s[..., 2] * s[..., 2])
    r = (F32(alpha) * F32(0.5)) * np.sqrt(n2)
    return np.minimum(np.maximum(r, F32(r_min)), F32(r_max)).astype(np.float32)


# --------------------------------------------------------------------------- a5
def grouping_operation(features, idx):
    """a5: out[b,c,j,s] = features[b,c,idx[b,j,s]].  (B,C,N),(B,P,S)->(B,C,P,S)."""
    features = np.asarray(features)
    idx = np.asarray(idx)
    B = features.shape[0]
    return np.stack([features[b][:, idx[b]] for b in range(B)], axis=0)


def grouping_operation_grad(grad_out, idx, N):
    """a5 backward: grad_features[b,c,idx[b,j,s]] += grad_out[b,c,j,s]."""
    grad_out = np.asarray(grad_out)
    B, C = grad_out.shape[:2]
    g = np.zeros((B, C, N), dtype=grad_out.dtype)
    for b in range(B):
        np.add.at(g[b], (slice(None), idx[b].reshape(-1)), grad_out[b].reshape(C, -1))
    return g


# --------------------------------------------------------------------------- a8
def three_nn(unknown, known):
    """a8.  unknown (B,n,3), known (B,m,3) -> dist (B,n,3) f32 = sqrt(d2) ascending,
    idx (B,n,3) i32.  Scan k ascending, insert on strict '<' == stable sort by
    (d2, k): ties go to the lowest index.  Requires m >= 3."""
    unknown = _f32c(unknown)
    known = _f32c(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    if m < 3:
        raise ValueError("three_nn requires m >= 3 known points")
    dist = np.zeros((B, n, 3), dtype=np.float32)
    idx = np.zeros((B, n, 3), dtype=np.int32)
    for b in range(B):
        d2 = sqdist(known[b][None, :, :], unknown[b][:, None, :])      # (n,m)
        order = np.argsort(d2, axis=1, kind="stable")[:, :3]
        idx[b] = order.astype(np.int32)
        dist[b] = np.sqrt(np.take_along_axis(d2, order, axis=1))
    return dist, idx


def interpolation_weights(dist):
    """FP-module weights: w = 1/(dist+1e-8), normalised by ((w0+w1)+w2)."""
    dist = _f32c(dist)
    recip = F32(1.0) / (dist + F32(1e-8))
    norm = (recip[..., 0] + recip[..., 1]) + recip[..., 2]
    return (recip / norm[..., None]).astype(np.float32)


# --------------------------------------------------------------------------- a9
def three_interpolate(features, idx, weight):
    """a9: out[b,c,i] = ((w0*f[idx0]) + (w1*f[idx1])) + (w2*f[idx2]), fp32 no FMA."""
    features = _f32c(features)
    weight = _f32c(weight)
    B, C, m = features.shape
    n = idx.shape[1]
    out = np.zeros((B, C, n), dtype=np.float32)
    for b in range(B):
        f = features[b]
        t0 = f[:, idx[b, :, 0]] * weight[b, :, 0][None, :]
        t1 = f[:, idx[b, :, 1]] * weight[b, :, 1][None, :]
        t2 = f[:, idx[b, :, 2]] * weight[b, :, 2][None, :]
        out[b] = (t0 + t1) + t2
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """a9 backward: grad_features[b,c,idx[b,i,t]] += grad_out[b,c,i]*weight[b,i,t]."""
    grad_out = _f32c(grad_out)
    weight = _f32c(weight)
    B, C, n = grad_out.shape
    g = np.zeros((B, C, m), dtype=np.float32)
    for b in range(B):
        for t in range(3):
            np.add.at(g[b], (slice(None), idx[b, :, t]), grad_out[b] * weight[b, :, t][None, :])
    return g


# --------------------------------------------------------------------------- a6
def bf16_round(a):
    """Round-to-nearest-even fp32 -> bf16 -> fp32 (used only for tight bf16 checks)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32).reshape(a.shape)


def shared_mlp(x, layers, pool=True, last_relu=True, emulate_bf16=False):
    """a6.  x (B,Cin,P,S) f32; layers = [(W (Cout,Cin), b (Cout,)), ...] with BN folded.
    y = ReLU(W.x + b) per layer (last ReLU optional), max over S if pool.
    -> (B,Cout,P) if pool else (B,Cout,P,S).  fp32 matmul (float64-free).
    emulate_bf16 rounds inputs, weights and inter-layer activations to bf16 with
    fp32 accumulation, mirroring the tcgen05 kernel's storage precision."""
    x = _f32c(x)
    B, Cin, P, S = x.shape
    h = x.transpose(0, 2, 3, 1).reshape(B * P * S, Cin)
    nl = len(layers)
    for li, (W, b) in enumerate(layers):
        W = _f32c(W)
        b = _f32c(b)
        if emulate_bf16:
            h = bf16_round(h)
            W = bf16_round(W)
        h = h @ W.T + b[None, :]
        if li < nl - 1 or last_relu:
            h = np.maximum(h, F32(0))
        h = h.astype(np.float32)
    Cout = h.shape[1]
    h = h.reshape(B, P, S, Cout)
    if pool:
        return np.ascontiguousarray(h.max(axis=2).transpose(0, 2, 1))
    return np.ascontiguousarray(h.transpose(0, 3, 1, 2))


# ---------------------------------------------------------------- modules (a7)
def _impl(impl):
    """Op provider for the module-level compositions: this file (NumPy) by default, or
    oracle.c_port (same results bit for bit, multi-threaded) for full-size / timed runs."""
    return impl if impl is not None else sys.modules[__name__]


def query_and_group(xyz, new_xyz, features, idx, radius, use_xyz=True, normalize_xyz=True, impl=None):
    """QueryAndGroup [LINEAGE]: grouped_xyz = xyz[idx] - new_xyz (/ radius if
    normalising; radius scalar or (B,npoint)); concat with grouped features on the
    channel axis -> (B, 3+C, npoint, nsample)."""
    xyz = _f32c(xyz)
    new_xyz = _f32c(new_xyz)
    B = xyz.shape[0]
    g = np.stack([xyz[b][idx[b]] for b in range(B)], axis=0)           # (B,P,S,3)
    g = g - new_xyz[:, :, None, :]
    if normalize_xyz:
        r = np.asarray(radius, dtype=np.float32)
        if r.ndim == 0:
            g = g / r
        else:
            g = g / r[:, :, None, None]
    g = np.ascontiguousarray(g.transpose(0, 3, 1, 2)).astype(np.float32)
    if features is None:
        return g
    gf = _impl(impl).grouping_operation(_f32c(features), idx)
    return np.concatenate([g, gf], axis=1) if use_xyz else gf


def sa_module(xyz, features, npoint, radius, nsample, layers, use_xyz=True,
              normalize_xyz=True, radius_t=None, inds=None, emulate_bf16=False, impl=None):
    """Set-abstraction module (SURVEY section 3 call stack 1).  radius_t (B,npoint)
    switches to the adaptive ball query and per-cluster normalisation.
    -> (new_xyz (B,npoint,3), new_features (B,Cout,npoint), inds (B,npoint))."""
    xyz = _f32c(xyz)
    B = xyz.shape[0]
    I = _impl(impl)
    if inds is None:
        inds = I.furthest_point_sample(xyz, npoint)
    new_xyz = np.stack([xyz[b][inds[b]] for b in range(B)], axis=0)
    if radius_t is not None:
        idx = I.ball_query_adaptive(radius_t, nsample, xyz, new_xyz)
        rr = _f32c(radius_t)
    else:
        idx = I.ball_query(radius, nsample, xyz, new_xyz)
        rr = F32(radius)
    x = query_and_group(xyz, new_xyz, features, idx, rr, use_xyz, normalize_xyz, impl=impl)
    out = shared_mlp(x, layers, pool=True, emulate_bf16=emulate_bf16)
    return new_xyz, out, inds


def fp_module(unknown, known, unknown_feats, known_feats, layers, emulate_bf16=False, impl=None):
    """Feature-propagation module (call stack 2) -> (B,Cout,n)."""
    I = _impl(impl)
    dist, idx = I.three_nn(unknown, known)
    w = interpolation_weights(dist)
    interp = I.three_interpolate(known_feats, idx, w)
    x = interp if unknown_feats is None else np.concatenate([interp, _f32c(unknown_feats)], axis=1)
    y = shared_mlp(x[..., None], layers, pool=False, emulate_bf16=emulate_bf16)
    return y[..., 0]


def voting_module(seed_xyz, seed_features, layers, emulate_bf16=False):
    """Voting glue [LINEAGE]: 3-layer 1x1 MLP (last layer linear) -> offsets (3) and
    feature residual (C).  vote_xyz = seed_xyz + offset; vote_features = seed + residual."""
    y = shared_mlp(_f32c(seed_features)[..., None], layers, pool=False, last_relu=False,
                   emulate_bf16=emulate_bf16)[..., 0]                 # (B,3+C,n)
    offset = y[:, :3, :].transpose(0, 2, 1)
    vote_xyz = (_f32c(seed_xyz) + offset).astype(np.float32)
    vote_features = (_f32c(seed_features) + y[:, 3:, :]).astype(np.float32)
    return vote_xyz, vote_features


def size_head(center_features, layers, scale=1.0, clip=2.0, emulate_bf16=False):
    """Size head [DECISION, unpinned]: 2-layer 1x1 MLP (last layer linear) on the vote features at
    the cluster centres (B,C,K) -> log-size y (B,3,K); size = scale * exp(clip(y, -clip, clip)) -> (B,K,3) f32."""
    y = shared_mlp(_f32c(center_features)[..., None], layers, pool=False, last_relu=False, emulate_bf16=emulate_bf16)[..., 0]
    y = np.minimum(np.maximum(y, F32(-clip)), F32(clip)).astype(np.float32)
    return np.ascontiguousarray((F32(scale) * np.exp(y)).astype(np.float32).transpose(0, 2, 1))


def vote_aggregation(vote_xyz, vote_features, size, npoint, nsample, layers,
                     alpha=1.0, r_min=0.1, r_max=1.2, emulate_bf16=False, impl=None):
    """a7: FPS over votes -> cluster centres -> per-cluster radius from predicted size
    (B,npoint,3) -> adaptive ball query -> group -> MLP + max-pool."""
    radius_t = size_to_radius(size, alpha, r_min, r_max)
    return sa_module(vote_xyz, vote_features, npoint, None, nsample, layers,
                     radius_t=radius_t, emulate_bf16=emulate_bf16, impl=impl) + (radius_t,)


def backbone_forward(xyz, features, params, cfg, emulate_bf16=False, impl=None):
    """4 SA + 2 FP backbone (call stack 3, SURVEY section 8 layer table).
    params: dict name -> layers; cfg: dict name -> (npoint, radius, nsample)."""
    end = {}
    x, f = _f32c(xyz), features
    for name in ("sa1", "sa2", "sa3", "sa4"):
        npoint, radius, nsample = cfg[name]
        x, f, inds = sa_module(x, f, npoint, radius, nsample, params[name], emulate_bf16=emulate_bf16,
                               impl=impl)
        end[name + "_xyz"], end[name + "_features"], end[name + "_inds"] = x, f, inds
    f = fp_module(end["sa3_xyz"], end["sa4_xyz"], end["sa3_features"], end["sa4_features"],
                  params["fp1"], emulate_bf16, impl)
    f = fp_module(end["sa2_xyz"], end["sa3_xyz"], end["sa2_features"], f, params["fp2"], emulate_bf16, impl)
    end["fp2_features"] = f
    end["fp2_xyz"] = end["sa2_xyz"]
    end["fp2_inds"] = end["sa1_inds"][:, : end["sa2_xyz"].shape[1]]
    return end


def detector_hot_path(xyz, features, size, params, cfg, emulate_bf16=False, impl=None):
    """The timed unit 'scene': backbone -> voting -> size-adaptive vote aggregation."""
    end = backbone_forward(xyz, features, params, cfg, emulate_bf16, impl)
    vxyz, vfeat = voting_module(end["fp2_xyz"], end["fp2_features"], params["vote"], emulate_bf16)
    npoint, _, nsample = cfg["agg"]
    cxyz, cfeat, cinds, radius_t = vote_aggregation(
        vxyz, vfeat, size, npoint, nsample, params["agg"],
        alpha=cfg.get("alpha", 1.0), r_min=cfg.get("r_min", 0.1), r_max=cfg.get("r_max", 1.2),
        emulate_bf16=emulate_bf16, impl=impl)
    end.update(vote_xyz=vxyz, vote_features=vfeat, cluster_xyz=cxyz, cluster_features=cfeat,
               cluster_inds=cinds, cluster_radius=radius_t)
    return end
