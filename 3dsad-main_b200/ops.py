"""PointNet++-lineage operator surface over libsad_b200.so.

Seven ``torch.autograd.Function`` subclasses exposed as callables with the lineage
argument orders (the mounted reference holds no code to cite -- README.md:1-2 only;
the rows cited are SURVEY.md section 8(a) a1..a9):

    furthest_point_sample(xyz, npoint)                    -> idx  (B,npoint)          i32
    gather_operation(features, idx)                       -> (B,C,npoint)             f32
    ball_query(radius, nsample, xyz, new_xyz)             -> idx  (B,npoint,nsample)  i32
    ball_query_adaptive(radius_t, nsample, xyz, new_xyz)  -> idx  (B,npoint,nsample)  i32
    grouping_operation(features, idx)                     -> (B,C,npoint,nsample)     f32
    three_nn(unknown, known)                              -> dist (B,n,3) f32, idx (B,n,3) i32
    three_interpolate(features, idx, weight)              -> (B,C,n)                  f32

Index outputs are int32 and non-differentiable.  Inputs must be CUDA, contiguous and
fp32 / int32; anything else raises (no CPU fallback, no silent copies).
"""
from __future__ import annotations

import ctypes

import torch
from torch.autograd import Function

from . import _lib


def _stream(t: torch.Tensor):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _p(t: torch.Tensor):
    return ctypes.c_void_p(t.data_ptr())


_VP0 = ctypes.c_void_p(0)


def _req(t, name, dtype, ndim, last=None):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: CUDA tensor required (libsad_b200 has no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if t.dim() != ndim:
        raise ValueError(f"{name}: expected {ndim} dims, got shape {tuple(t.shape)}")
    if last is not None and t.shape[-1] != last:
        raise ValueError(f"{name}: last dim must be {last}, got shape {tuple(t.shape)}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous (call .contiguous() first)")
    return t


def _same_dev(*ts):
    d = ts[0].device
    for t in ts[1:]:
        if t.device != d:
            raise RuntimeError("all tensors must live on the same CUDA device")
    return d


class SceneGrid:
    """Spatial sort of a batch of scenes (csrc/grid.cu): points ordered by cell of a 32^3 grid plus the
    cell offsets.  Built once per coordinate tensor and shared by the culled FPS and the grid ball
    query over the same `xyz`; results of both are bit-identical to the plain kernels."""

    __slots__ = ("xyz", "B", "N", "workspace", "version")

    def __init__(self, xyz: torch.Tensor):
        _req(xyz, "xyz", torch.float32, 3, 3)
        self.xyz, (self.B, self.N) = xyz, xyz.shape[:2]
        self.version = xyz._version
        lib = _lib.load()
        nbytes = int(lib.sad_scene_grid_workspace_bytes(self.B, self.N))
        if nbytes < 0:
            raise ValueError("scene grid: bad sizes")
        self.workspace = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=xyz.device)
        with torch.cuda.device(xyz.device):
            _lib.check(lib.sad_scene_grid_build(self.B, self.N, _p(xyz), _p(self.workspace), _stream(xyz)),
                       "scene_grid_build")

    def check(self, xyz: torch.Tensor):
        """The grid holds its own sorted copy of the coordinates: it is valid for the tensor it was built from, as long
        as that tensor has not been written since (in-place updates bump `_version`).  A grid is single-stream: the
        culled FPS kernels keep their min-distance scratch in its workspace."""
        if xyz.data_ptr() != self.xyz.data_ptr() or tuple(xyz.shape) != tuple(self.xyz.shape):
            raise ValueError("SceneGrid was built for a different xyz tensor")
        if xyz._version != self.version:
            raise ValueError("SceneGrid is stale: xyz was modified in place after the grid was built (rebuild it with "
                             "build_scene_grid)")
        return self


def build_scene_grid(xyz: torch.Tensor) -> SceneGrid:
    return SceneGrid(xyz)


# Scenes at least this large get a scene grid built on the fly by furthest_point_sample / ball_query
# when the caller passes none (the build costs one short kernel; below it the plain kernels win).
GRID_MIN_POINTS = 8192

# Backward of gather / grouping / three_interpolate: False = fp32 atomics (fastest; the summation order, hence the last
# bits, vary from run to run), True = sort-by-destination plan + sequential per-destination sums (csrc/scatter.cu):
# reproducible bit for bit and equal to the oracle's index-order accumulation.
DETERMINISTIC = [False]


def set_deterministic(on: bool = True) -> bool:
    """Switch the scatter-add backward kernels to the deterministic mode; returns the previous setting."""
    prev = DETERMINISTIC[0]
    DETERMINISTIC[0] = bool(on)
    return prev


def _scatter_add_det(grad_out3, idx2, N):
    """grad_out3 (B,C,PS) f32, idx2 (B,PS) i32 -> (B,C,N) f32, deterministic."""
    B, C, PS = grad_out3.shape
    dev = grad_out3.device
    order = torch.empty((B, max(PS, 1)), dtype=torch.int32, device=dev)
    offsets = torch.empty((B, N + 1), dtype=torch.int32, device=dev)
    g = torch.empty((B, C, N), dtype=torch.float32, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.sad_scatter_plan_build(B, N, PS, _p(idx2), _p(order), _p(offsets), _stream(grad_out3)), "scatter_plan_build")
        _lib.check(lib.sad_scatter_add_det(B, C, N, PS, _p(grad_out3), _p(order), _p(offsets), _p(g), _stream(grad_out3)),
                   "scatter_add_det")
    return g
FPS_POLICIES = {"latency": 0, "throughput": 1, "throughput_paired": 2}   # include/sad_ops.h SAD_FPS_*


# Under torch.autocast the differentiable ops run in fp32 (bf16 features are cast on entry, gradients come back in the
# input's dtype): the point-wise MLPs are the only bf16 consumers of a mixed-precision training step (config 4).
_amp_fwd = torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_amp_bwd = torch.amp.custom_bwd(device_type="cuda")


class FurthestPointSampling(Function):
    """a1.  xyz (B,N,3) f32 -> (B,npoint) i32; sel[0]=0, ties to the lowest index.
    Optional `grid` (SceneGrid of the same xyz) selects the exact culled kernel; `policy` picks how it is scheduled
    ("latency": a cluster of SMs per scene, "throughput": one SM per scene, "throughput_paired": two scenes per
    SM) -- never the result.  `variant` (tests / tools): forced cluster size of the kernel, 0 = the default."""

    @staticmethod
    def forward(ctx, xyz, npoint, grid=None, policy="latency", prefix_ordered=False, variant=0):
        _req(xyz, "xyz", torch.float32, 3, 3)
        B, N, _ = xyz.shape
        npoint = int(npoint)
        if npoint < 1 or N < 1:
            raise ValueError("furthest_point_sample: need N >= 1 and npoint >= 1")
        lib = _lib.load()
        out = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
        if prefix_ordered and npoint <= N:
            # xyz is in farthest-point order already (the previous stage's new_xyz): identity behind a device-side
            # duplicate guard, the sampler only for scenes that fail it -- same result either way
            flags = torch.empty((B,), dtype=torch.int32, device=xyz.device)
            with torch.cuda.device(xyz.device):
                _lib.check(lib.sad_furthest_point_sample_prefix_fwd(B, N, npoint, _p(xyz), _p(out), _p(flags),
                                                                    _stream(xyz)), "furthest_point_sample_prefix")
            ctx.mark_non_differentiable(out)
            return out
        if grid is None and GRID_MIN_POINTS <= N <= lib.sad_fps_grid_max_points():
            grid = SceneGrid(xyz)
        with torch.cuda.device(xyz.device):
            if grid is not None and N <= lib.sad_fps_grid_max_points():
                grid.check(xyz)
                _lib.check(lib.sad_furthest_point_sample_grid_policy_fwd(
                    B, N, npoint, _p(xyz), _p(grid.workspace), _p(out), FPS_POLICIES[policy], int(variant), _stream(xyz)),
                    "furthest_point_sample_grid")
            else:
                _lib.check(lib.sad_furthest_point_sample_cs_fwd(B, N, npoint, _p(xyz), _p(out), int(variant), _stream(xyz)),
                           "furthest_point_sample")
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad=None):
        return None, None, None, None, None, None


class GatherOperation(Function):
    """a2.  features (B,C,N) f32, idx (B,npoint) i32 -> (B,C,npoint)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, features, idx):
        _req(features, "features", torch.float32, 3)
        _req(idx, "idx", torch.int32, 2)
        _same_dev(features, idx)
        B, C, N = features.shape
        if idx.shape[0] != B:
            raise ValueError("gather_operation: batch mismatch")
        npoint = idx.shape[1]
        out = torch.empty((B, C, npoint), dtype=torch.float32, device=features.device)
        with torch.cuda.device(features.device):
            _lib.check(_lib.load().sad_gather_operation_fwd(B, C, N, npoint, _p(features), _p(idx), _p(out),
                                                            _stream(features)), "gather_operation")
        ctx.save_for_backward(idx)
        ctx.N = N
        return out

    @staticmethod
    @_amp_bwd
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        B, C, npoint = grad_out.shape
        if DETERMINISTIC[0]:
            return _scatter_add_det(grad_out, idx, ctx.N), None
        g = torch.empty((B, C, ctx.N), dtype=torch.float32, device=grad_out.device)
        with torch.cuda.device(grad_out.device):
            _lib.check(_lib.load().sad_gather_operation_bwd(B, C, ctx.N, npoint, _p(grad_out), _p(idx), _p(g),
                                                            _stream(grad_out)), "gather_operation_bwd")
        return g, None


class BallQuery(Function):
    """a3.  (radius, nsample, xyz (B,N,3), new_xyz (B,npoint,3)) -> idx (B,npoint,nsample) i32.
    Optional `grid` (SceneGrid of the same xyz) selects the grid-accelerated kernel."""

    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz, grid=None):
        _req(xyz, "xyz", torch.float32, 3, 3)
        _req(new_xyz, "new_xyz", torch.float32, 3, 3)
        _same_dev(xyz, new_xyz)
        B, N, _ = xyz.shape
        if new_xyz.shape[0] != B:
            raise ValueError("ball_query: batch mismatch")
        npoint, nsample = new_xyz.shape[1], int(nsample)
        if nsample < 1:
            raise ValueError("ball_query: nsample must be >= 1")
        out = torch.empty((B, npoint, nsample), dtype=torch.int32, device=xyz.device)
        if grid is None and N >= GRID_MIN_POINTS:
            grid = SceneGrid(xyz)
        with torch.cuda.device(xyz.device):
            if grid is not None:
                grid.check(xyz)
                _lib.check(_lib.load().sad_ball_query_grid_fwd(B, N, npoint, float(radius), _VP0, nsample, _p(xyz),
                                                               _p(grid.workspace), _p(new_xyz), _p(out), _stream(xyz)),
                           "ball_query_grid")
            else:
                _lib.check(_lib.load().sad_ball_query_fwd(B, N, npoint, float(radius), nsample, _p(xyz), _p(new_xyz),
                                                          _p(out), _stream(xyz)), "ball_query")
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad=None):
        return None, None, None, None, None


class BallQueryAdaptive(Function):
    """a4 (3DSAD).  radius_t (B,npoint) f32: per-cluster radius from predicted object size.
    Optional `grid` as for ball_query."""

    @staticmethod
    def forward(ctx, radius_t, nsample, xyz, new_xyz, grid=None):
        _req(xyz, "xyz", torch.float32, 3, 3)
        _req(new_xyz, "new_xyz", torch.float32, 3, 3)
        _req(radius_t, "radius_t", torch.float32, 2)
        _same_dev(xyz, new_xyz, radius_t)
        B, N, _ = xyz.shape
        npoint, nsample = new_xyz.shape[1], int(nsample)
        if new_xyz.shape[0] != B or tuple(radius_t.shape) != (B, npoint):
            raise ValueError("ball_query_adaptive: radius_t must have shape (B, npoint)")
        if nsample < 1:
            raise ValueError("ball_query_adaptive: nsample must be >= 1")
        out = torch.empty((B, npoint, nsample), dtype=torch.int32, device=xyz.device)
        if grid is None and N >= GRID_MIN_POINTS:
            grid = SceneGrid(xyz)
        with torch.cuda.device(xyz.device):
            if grid is not None:
                grid.check(xyz)
                _lib.check(_lib.load().sad_ball_query_grid_fwd(B, N, npoint, 0.0, _p(radius_t), nsample, _p(xyz),
                                                               _p(grid.workspace), _p(new_xyz), _p(out), _stream(xyz)),
                           "ball_query_grid")
            else:
                _lib.check(_lib.load().sad_ball_query_adaptive_fwd(B, N, npoint, _p(radius_t), nsample, _p(xyz),
                                                                   _p(new_xyz), _p(out), _stream(xyz)),
                           "ball_query_adaptive")
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad=None):
        return None, None, None, None, None


class GroupingOperation(Function):
    """a5.  features (B,C,N) f32, idx (B,npoint,nsample) i32 -> (B,C,npoint,nsample)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, features, idx):
        _req(features, "features", torch.float32, 3)
        _req(idx, "idx", torch.int32, 3)
        _same_dev(features, idx)
        B, C, N = features.shape
        if idx.shape[0] != B:
            raise ValueError("grouping_operation: batch mismatch")
        _, npoint, nsample = idx.shape
        out = torch.empty((B, C, npoint, nsample), dtype=torch.float32, device=features.device)
        with torch.cuda.device(features.device):
            _lib.check(_lib.load().sad_grouping_operation_fwd(B, C, N, npoint, nsample, _p(features), _p(idx),
                                                              _p(out), _stream(features)), "grouping_operation")
        ctx.save_for_backward(idx)
        ctx.N = N
        return out

    @staticmethod
    @_amp_bwd
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        B, C, npoint, nsample = grad_out.shape
        if DETERMINISTIC[0]:
            return _scatter_add_det(grad_out.view(B, C, npoint * nsample), idx.view(B, npoint * nsample), ctx.N), None
        g = torch.empty((B, C, ctx.N), dtype=torch.float32, device=grad_out.device)
        with torch.cuda.device(grad_out.device):
            _lib.check(_lib.load().sad_grouping_operation_bwd(B, C, ctx.N, npoint, nsample, _p(grad_out), _p(idx),
                                                              _p(g), _stream(grad_out)), "grouping_operation_bwd")
        return g, None


class ThreeNN(Function):
    """a8.  unknown (B,n,3), known (B,m,3), m >= 3 -> dist (B,n,3) f32, idx (B,n,3) i32."""

    @staticmethod
    def forward(ctx, unknown, known):
        _req(unknown, "unknown", torch.float32, 3, 3)
        _req(known, "known", torch.float32, 3, 3)
        _same_dev(unknown, known)
        B, n, _ = unknown.shape
        m = known.shape[1]
        if known.shape[0] != B:
            raise ValueError("three_nn: batch mismatch")
        if m < 3:
            raise ValueError("three_nn requires m >= 3 known points")
        dist = torch.empty((B, n, 3), dtype=torch.float32, device=unknown.device)
        idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknown.device)
        with torch.cuda.device(unknown.device):
            _lib.check(_lib.load().sad_three_nn_fwd(B, n, m, _p(unknown), _p(known), _p(dist), _p(idx),
                                                    _stream(unknown)), "three_nn")
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, a=None, b=None):
        return None, None


class ThreeInterpolate(Function):
    """a9.  features (B,C,m) f32, idx (B,n,3) i32, weight (B,n,3) f32 -> (B,C,n)."""

    @staticmethod
    @_amp_fwd
    def forward(ctx, features, idx, weight):
        _req(features, "features", torch.float32, 3)
        _req(idx, "idx", torch.int32, 3, 3)
        _req(weight, "weight", torch.float32, 3, 3)
        _same_dev(features, idx, weight)
        B, C, m = features.shape
        n = idx.shape[1]
        if idx.shape[0] != B or tuple(weight.shape) != tuple(idx.shape):
            raise ValueError("three_interpolate: idx / weight must both be (B,n,3)")
        out = torch.empty((B, C, n), dtype=torch.float32, device=features.device)
        with torch.cuda.device(features.device):
            _lib.check(_lib.load().sad_three_interpolate_fwd(B, C, m, n, _p(features), _p(idx), _p(weight),
                                                             _p(out), _stream(features)), "three_interpolate")
        ctx.save_for_backward(idx, weight)
        ctx.m = m
        return out

    @staticmethod
    @_amp_bwd
    def backward(ctx, grad_out):
        idx, weight = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        B, C, n = grad_out.shape
        g = torch.empty((B, C, ctx.m), dtype=torch.float32, device=grad_out.device)
        if DETERMINISTIC[0]:
            order = torch.empty((B, max(3 * n, 1)), dtype=torch.int32, device=grad_out.device)
            offsets = torch.empty((B, ctx.m + 1), dtype=torch.int32, device=grad_out.device)
            lib = _lib.load()
            with torch.cuda.device(grad_out.device):
                _lib.check(lib.sad_interp_plan_build(B, n, ctx.m, _p(idx), _p(order), _p(offsets), _stream(grad_out)),
                           "interp_plan_build")
                _lib.check(lib.sad_three_interpolate_bwd_det(B, C, n, ctx.m, _p(grad_out), _p(weight), _p(order), _p(offsets),
                                                             _p(g), _stream(grad_out)), "three_interpolate_bwd_det")
            return g, None, None
        with torch.cuda.device(grad_out.device):
            _lib.check(_lib.load().sad_three_interpolate_bwd(B, C, n, ctx.m, _p(grad_out), _p(idx), _p(weight),
                                                             _p(g), _stream(grad_out)), "three_interpolate_bwd")
        return g, None, None


furthest_point_sample = FurthestPointSampling.apply
gather_operation = GatherOperation.apply
ball_query = BallQuery.apply
ball_query_adaptive = BallQueryAdaptive.apply
grouping_operation = GroupingOperation.apply
three_nn = ThreeNN.apply
three_interpolate = ThreeInterpolate.apply


def size_to_radius(size: torch.Tensor, alpha: float = 1.0, r_min: float = 0.1, r_max: float = 1.2):
    """Predicted box size (B,K,3) -> per-cluster radius (B,K):
    r = clamp(alpha * 0.5 * ||size||_2, r_min, r_max)   [SURVEY a4, DECISION: formula unpinned].
    One kernel on CUDA inputs without autograd (bit-equal to the torch expression below and to the oracle)."""
    if size.is_cuda and not (torch.is_grad_enabled() and size.requires_grad) and size.dtype == torch.float32 \
            and size.dim() >= 1 and size.shape[-1] == 3:
        s = size.contiguous()
        out = torch.empty(s.shape[:-1], dtype=torch.float32, device=s.device)
        with torch.cuda.device(s.device):
            _lib.check(_lib.load().sad_size_to_radius(out.numel(), _p(s), float(alpha), float(r_min), float(r_max), _p(out),
                                                      _stream(s)), "size_to_radius")
        return out
    s = size.float()
    n2 = (s[..., 0] * s[..., 0] + s[..., 1] * s[..., 1]) + s[..., 2] * s[..., 2]
    r = (alpha * 0.5) * torch.sqrt(n2)
    return r.clamp(min=r_min, max=r_max).contiguous()


def gather_points(xyz: torch.Tensor, inds: torch.Tensor, with_xyzw: bool = False):
    """new_xyz (B,npoint,3) = xyz[b, inds[b, j], :] in one launch, no autograd (coordinates are inputs).  With
    `with_xyzw` also returns the padded (B,npoint,4) copy the fused SA kernel gathers its special K step from."""
    _req(xyz, "xyz", torch.float32, 3, 3)
    _req(inds, "inds", torch.int32, 2)
    _same_dev(xyz, inds)
    B, N, _ = xyz.shape
    P = inds.shape[1]
    out = torch.empty((B, P, 3), dtype=torch.float32, device=xyz.device)
    xyzw = torch.empty((B, P, 4), dtype=torch.float32, device=xyz.device) if with_xyzw else None
    with torch.cuda.device(xyz.device):
        _lib.check(_lib.load().sad_gather_points_fwd(B, N, P, _p(xyz), _p(inds), _p(out),
                                                     _p(xyzw) if xyzw is not None else _VP0, _stream(xyz)), "gather_points")
    return (out, xyzw) if with_xyzw else out


def three_nn_weights(unknown: torch.Tensor, known: torch.Tensor):
    """three_nn plus the FP module's normalised inverse-distance weights in the same launch
    -> dist (B,n,3), idx (B,n,3) i32, weight (B,n,3); no autograd (coordinates only)."""
    _req(unknown, "unknown", torch.float32, 3, 3)
    _req(known, "known", torch.float32, 3, 3)
    _same_dev(unknown, known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    if known.shape[0] != B:
        raise ValueError("three_nn: batch mismatch")
    if m < 3:
        raise ValueError("three_nn requires m >= 3 known points")
    dist = torch.empty((B, n, 3), dtype=torch.float32, device=unknown.device)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknown.device)
    weight = torch.empty((B, n, 3), dtype=torch.float32, device=unknown.device)
    with torch.cuda.device(unknown.device):
        _lib.check(_lib.load().sad_three_nn_weights_fwd(B, n, m, _p(unknown), _p(known), _p(dist), _p(idx), _p(weight),
                                                        _stream(unknown)), "three_nn_weights")
    return dist, idx, weight
