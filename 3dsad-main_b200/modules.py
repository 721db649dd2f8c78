"""SA / FP / voting / size-adaptive vote-aggregation modules and the backbone built from
the operator surface in ops.py (SURVEY.md section 3 call stacks 1-3, section 8(b)).

Module signatures follow the PointNet++ / VoteNet lineage that BASELINE.json's
north_star names (the mounted reference has no code to cite, README.md:1-2 only):

    PointnetSAModuleVotes(npoint, radius, nsample, mlp, use_xyz=True, normalize_xyz=True)
        forward(xyz (B,N,3), features (B,C,N), inds=None, radius_t=None)
            -> new_xyz (B,npoint,3), new_features (B,C_out,npoint), inds (B,npoint)
    PointnetFPModule(mlp)
        forward(unknown (B,n,3), known (B,m,3), unknow_feats (B,C1,n), known_feats (B,C2,m))
            -> (B,C_out,n)

Two execution paths, same results within the bf16 tolerance:
  * eval (inference, the timed path): BN folded into the 1x1 convs, the shared MLP runs
    through sad_b200.mlp (hand-written tcgen05 kernels);
  * train: plain torch Conv2d/BatchNorm2d/ReLU so autograd + DDP work; the point ops'
    backward kernels (scatter-add) come from ops.py.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops
from . import mlp as _mlp
from .config import LAYER_CFG, mlp_channels


# How the scene-grid FPS of large scenes is scheduled (ops.FurthestPointSampling `policy`): "latency" = a cluster of SMs
# per scene (an eager caller waiting for one batch), "throughput" = one SM per scene (engine.PipelinedHotPath switches
# to it while it captures its graphs: with several batches in flight the step is bound by SM-time, not by any one
# kernel's latency).  Results are bit-identical.
FPS_POLICY = ["latency"]


class SharedMLP(nn.Module):
    """Stack of 1x1 Conv2d (+BatchNorm2d) + ReLU over (B,C,P,S) [LINEAGE pt_utils.SharedMLP]."""

    def __init__(self, channels: Sequence[int], bn: bool = True, last_relu: bool = True):
        super().__init__()
        self.channels = list(channels)
        self.last_relu = last_relu
        self.convs = nn.ModuleList()
        self.bns = nn.ModuleList()
        for cin, cout in zip(channels[:-1], channels[1:]):
            self.convs.append(nn.Conv2d(cin, cout, kernel_size=1, bias=not bn))
            self.bns.append(nn.BatchNorm2d(cout) if bn else nn.Identity())
        self._folded = None
        self._folded_key = None
        self.mlp_dtype = "bf16"      # operand precision of the fused inference kernels: "bf16" | "tf32" (mlp.PreparedMLP)

    def train(self, mode: bool = True):
        self._folded = None
        return super().train(mode)

    def _fold_key(self):
        """Identity of everything the folded weights depend on: device / dtype moves, load_state_dict, optimizer or
        in-place updates and BN-statistic changes all change a data pointer or a version counter."""
        key = []
        for conv, bn in zip(self.convs, self.bns):
            ts = [conv.weight, conv.bias]
            if isinstance(bn, nn.BatchNorm2d):
                ts += [bn.weight, bn.bias, bn.running_mean, bn.running_var]
            for t in ts:
                key.append(None if t is None else (t.data_ptr(), t._version, t.device, t.dtype))
        key.append(self.mlp_dtype)
        return tuple(key)

    @torch.no_grad()
    def load_folded(self, layers):
        """Install [(W (Cout,Cin), b (Cout,)), ...] (numpy or tensors) as conv weights with identity BN."""
        for conv, bn, (W, b) in zip(self.convs, self.bns, layers):
            W = torch.as_tensor(W, dtype=torch.float32)
            b = torch.as_tensor(b, dtype=torch.float32)
            conv.weight.copy_(W.view(*W.shape, 1, 1))
            if isinstance(bn, nn.BatchNorm2d):
                bn.running_mean.zero_()
                bn.running_var.fill_(1.0)
                bn.weight.fill_(float((1.0 + bn.eps) ** 0.5))
                bn.bias.copy_(b)
            else:
                conv.bias.copy_(b)
        self._folded = None

    @torch.no_grad()
    def folded(self) -> List[Tuple[torch.Tensor, torch.Tensor]]:
        """Eval-mode BN folded into (W, b) per layer.  Cached, and rebuilt whenever a parameter / buffer it was
        derived from has changed (see _fold_key); graphs captured by engine.PipelinedHotPath bake the weights of
        capture time in and must be re-captured after a weight update."""
        key = self._fold_key()
        if self._folded is None or self._folded_key != key:
            out = []
            for conv, bn in zip(self.convs, self.bns):
                W = conv.weight.detach().flatten(1).float()
                b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(W.shape[0], device=W.device)
                if isinstance(bn, nn.BatchNorm2d):
                    s = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
                    W = W * s[:, None]
                    b = (b - bn.running_mean) * s + bn.bias.detach()
                out.append((W.contiguous(), b.contiguous()))
            self._folded = _mlp.prepare_layers(out, self.mlp_dtype)
            self._folded_key = key
        return self._folded

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n = len(self.convs)
        for i, (conv, bn) in enumerate(zip(self.convs, self.bns)):
            x = bn(conv(x))
            if i < n - 1 or self.last_relu:
                x = torch.relu(x)
        return x


def _gather_xyz(xyz: torch.Tensor, inds: torch.Tensor) -> torch.Tensor:
    """new_xyz (B,npoint,3) = xyz[b, inds[b]]: one launch when no gradient is wanted, else through gather_operation
    (lineage idiom: transpose, gather, transpose)."""
    if not (torch.is_grad_enabled() and xyz.requires_grad):
        out, xyzw = ops.gather_points(xyz, inds, with_xyzw=True)
        out._sad_xyzw = xyzw          # padded twin: what the next stage's fused kernel gathers its special K step from
        return out
    flipped = xyz.transpose(1, 2).contiguous()
    return ops.gather_operation(flipped, inds).transpose(1, 2).contiguous()


def query_and_group(xyz, new_xyz, features, idx, radius, use_xyz=True, normalize_xyz=True):
    """[LINEAGE QueryAndGroup] -> (B, 3+C, npoint, nsample); radius scalar or (B,npoint)."""
    grouped_xyz = ops.grouping_operation(xyz.transpose(1, 2).contiguous(), idx)
    grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
    if normalize_xyz:
        if torch.is_tensor(radius):
            grouped_xyz = grouped_xyz / radius[:, None, :, None]
        else:
            grouped_xyz = grouped_xyz / float(radius)
    if features is None:
        return grouped_xyz
    grouped = ops.grouping_operation(features, idx)
    return torch.cat([grouped_xyz, grouped], dim=1) if use_xyz else grouped


class PointnetSAModuleVotes(nn.Module):
    """Set abstraction: FPS -> (adaptive) ball query -> group -> shared MLP -> max-pool."""

    def __init__(self, npoint: int, radius: Optional[float], nsample: int, mlp: Sequence[int],
                 use_xyz: bool = True, normalize_xyz: bool = True, bn: bool = True):
        super().__init__()
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.use_xyz, self.normalize_xyz = use_xyz, normalize_xyz
        # set by a caller that feeds this stage the previous stage's new_xyz (farthest-point order): the sampling is
        # then the identity behind a device-side duplicate guard (ops.furthest_point_sample prefix_ordered)
        self.prefix_ordered_input = False
        ch = list(mlp)
        if use_xyz:
            ch[0] += 3
        self.mlp_module = SharedMLP(ch, bn=bn)

    def forward(self, xyz, features=None, inds=None, radius_t=None, new_xyz=None, grid=None):
        """`grid`: optional ops.SceneGrid of `xyz` shared by the sampling and the neighbour search (large
        scenes build one on the fly when none is passed)."""
        if grid is None and inds is None and xyz.shape[1] >= ops.GRID_MIN_POINTS:
            grid = ops.build_scene_grid(xyz)
        if inds is None:
            inds = ops.furthest_point_sample(xyz, self.npoint, grid, FPS_POLICY[0], self.prefix_ordered_input)
        if new_xyz is None:
            new_xyz = _gather_xyz(xyz, inds)
        if radius_t is not None:
            idx = ops.ball_query_adaptive(radius_t, self.nsample, xyz, new_xyz, grid)
            rad = radius_t
        else:
            idx = ops.ball_query(self.radius, self.nsample, xyz, new_xyz, grid)
            rad = self.radius
        if not self.training and not torch.is_grad_enabled():
            new_features = _mlp.sa_group_mlp(xyz, new_xyz, features, idx, rad, self.mlp_module.folded(),
                                             use_xyz=self.use_xyz, normalize_xyz=self.normalize_xyz)
        else:
            grouped = query_and_group(xyz, new_xyz, features, idx, rad, self.use_xyz, self.normalize_xyz)
            new_features = self.mlp_module(grouped).max(dim=3)[0]
        return new_xyz, new_features, inds


class PointnetFPModule(nn.Module):
    """Feature propagation: three_nn -> inverse-distance weights -> three_interpolate -> MLP."""

    def __init__(self, mlp: Sequence[int], bn: bool = True):
        super().__init__()
        self.mlp = SharedMLP(list(mlp), bn=bn)

    @staticmethod
    def interpolation_plan(unknown, known):
        """three_nn + inverse-distance weights (depends on coordinates only) -> (idx, weight)."""
        if not (torch.is_grad_enabled() and (unknown.requires_grad or known.requires_grad)):
            _, idx, weight = ops.three_nn_weights(unknown, known)      # one launch, weights bit-equal to the lines below
            return idx, weight
        dist, idx = ops.three_nn(unknown, known)
        dist_recip = 1.0 / (dist + 1e-8)
        norm = (dist_recip[..., 0] + dist_recip[..., 1]) + dist_recip[..., 2]
        return idx, (dist_recip / norm.unsqueeze(-1)).contiguous()

    def forward(self, unknown, known, unknow_feats, known_feats, plan=None):
        fast = not self.training and not torch.is_grad_enabled()
        if known is not None:
            idx, weight = plan if plan is not None else self.interpolation_plan(unknown, known)
            if fast:
                return _mlp.fp_interp_mlp(known_feats, unknow_feats, idx, weight, self.mlp.folded())
            interpolated = ops.three_interpolate(known_feats.contiguous(), idx, weight)
        else:
            interpolated = known_feats.expand(*known_feats.size()[0:2], unknown.size(1))
        new_features = interpolated if unknow_feats is None else torch.cat([interpolated, unknow_feats], dim=1)
        if not self.training and not torch.is_grad_enabled():
            return _mlp.pointwise_mlp(new_features, self.mlp.folded(), last_relu=True)
        return self.mlp(new_features.unsqueeze(-1)).squeeze(-1)


class Pointnet2Backbone(nn.Module):
    """4 SA + 2 FP backbone (SURVEY section 8 layer table) -> 1024 seeds x 256 features."""

    def __init__(self, input_feature_dim: int = 1, bn: bool = True):
        super().__init__()
        ch = mlp_channels(input_feature_dim)
        c = LAYER_CFG
        self.sa1 = PointnetSAModuleVotes(*c["sa1"], mlp=[input_feature_dim] + ch["sa1"][1:], bn=bn)
        self.sa2 = PointnetSAModuleVotes(*c["sa2"], mlp=[128] + ch["sa2"][1:], bn=bn)
        self.sa3 = PointnetSAModuleVotes(*c["sa3"], mlp=[256] + ch["sa3"][1:], bn=bn)
        self.sa4 = PointnetSAModuleVotes(*c["sa4"], mlp=[256] + ch["sa4"][1:], bn=bn)
        self.fp1 = PointnetFPModule(ch["fp1"], bn=bn)
        self.fp2 = PointnetFPModule(ch["fp2"], bn=bn)
        for m in (self.sa2, self.sa3, self.sa4):
            m.prefix_ordered_input = True

    overlap_geometry = True     # run the coordinate-only chain (FPS, three_nn) on a side stream

    def _side_streams(self, device):
        st = getattr(self, "_geo_streams", None)
        if st is None or st[0].device != device:
            st = self._geo_streams = (torch.cuda.Stream(device=device), torch.cuda.Stream(device=device))
        return st

    def _geometry_chain(self, xyz, ready_event=None):
        """Everything that depends on coordinates only, issued on two side streams:
          geoA : FPS over the raw scene (SA1) -- the long serial kernel;
          geoB : the three smaller dependent FPS passes and the two interpolation plans.
        The serial FPS latency then overlaps the ball queries / MLPs of the earlier stages, and --
        when the caller passes `ready_event` (inputs already resident) instead of ordering against
        the whole main stream -- the feature work of the PREVIOUS batch as well."""
        main = torch.cuda.current_stream(xyz.device)
        geo_a, geo_b = self._side_streams(xyz.device)
        if ready_event is not None:
            geo_a.wait_event(ready_event)
        else:
            geo_a.wait_stream(main)
        plan = {}
        xyz.record_stream(geo_a)
        with torch.cuda.stream(geo_a):
            grid = ops.build_scene_grid(xyz) if xyz.shape[1] >= ops.GRID_MIN_POINTS else None
            inds = ops.furthest_point_sample(xyz, self.sa1.npoint, grid, FPS_POLICY[0])
            x = _gather_xyz(xyz, inds)
            ev = torch.cuda.Event()
            ev.record(geo_a)
            plan["sa1"] = (inds, x, ev)
            plan["grid"] = grid
            if grid is not None:
                grid.workspace.record_stream(main)
        geo_b.wait_event(ev)
        with torch.cuda.stream(geo_b):
            for name in ("sa2", "sa3", "sa4"):
                # x is the previous stage's new_xyz, i.e. already in farthest-point order (SURVEY H3 side note)
                inds = ops.furthest_point_sample(x, getattr(self, name).npoint, None, "latency", True)
                x = _gather_xyz(x, inds)
                ev = torch.cuda.Event()
                ev.record(geo_b)
                plan[name] = (inds, x, ev)
            p1 = PointnetFPModule.interpolation_plan(plan["sa3"][1], plan["sa4"][1])
            p2 = PointnetFPModule.interpolation_plan(plan["sa2"][1], plan["sa3"][1])
            ev = torch.cuda.Event()
            ev.record(geo_b)
            plan["fp"] = (p1, p2, ev)
        plan["sa1"][1].record_stream(geo_b)
        for k, v in plan.items():         # tensors born on a side stream, consumed on `main`
            if k == "grid":
                continue
            for t in v:
                if torch.is_tensor(t):
                    t.record_stream(main)
                elif isinstance(t, tuple):
                    for u in t:
                        u.record_stream(main)
        return plan, main

    def forward(self, xyz, features, ready_event=None):
        end = {}
        x, f = xyz, features
        fast = (self.overlap_geometry and not self.training and not torch.is_grad_enabled() and xyz.is_cuda)
        plan = None
        if fast:
            plan, main = self._geometry_chain(xyz, ready_event)
            if features is not None and features.shape[1] == 1 and self.sa1.use_xyz:
                _mlp.prepack_xyzw(xyz, features)      # SA1's gathered source, built on the main stream under the sampling chain
        for name in ("sa1", "sa2", "sa3", "sa4"):
            if plan is not None:
                inds, new_xyz, ev = plan[name]
                main.wait_event(ev)
                x, f, inds = getattr(self, name)(x, f, inds=inds, new_xyz=new_xyz,
                                                 grid=plan["grid"] if name == "sa1" else None)
            else:
                x, f, inds = getattr(self, name)(x, f)
            end[name + "_xyz"], end[name + "_features"], end[name + "_inds"] = x, f, inds
        p1 = p2 = None
        if plan is not None:
            p1, p2, ev = plan["fp"]
            main.wait_event(ev)
        f = self.fp1(end["sa3_xyz"], end["sa4_xyz"], end["sa3_features"], end["sa4_features"], plan=p1)
        f = self.fp2(end["sa2_xyz"], end["sa3_xyz"], end["sa2_features"], f, plan=p2)
        end["fp2_features"] = f
        end["fp2_xyz"] = end["sa2_xyz"]
        end["fp2_inds"] = end["sa1_inds"][:, : end["sa2_xyz"].shape[1]]
        return end


class VotingModule(nn.Module):
    """Seed -> vote glue [LINEAGE]: 3-layer 1x1 MLP, last layer linear; vote = seed + offset."""

    def __init__(self, seed_feature_dim: int = 256, bn: bool = True):
        super().__init__()
        c = seed_feature_dim
        self.mlp = SharedMLP([c, c, c, 3 + c], bn=bn, last_relu=False)
        self.mlp.bns[-1] = nn.Identity()          # last conv is a plain linear layer with bias
        self.mlp.convs[-1] = nn.Conv2d(c, 3 + c, kernel_size=1, bias=True)

    def forward(self, seed_xyz, seed_features):
        if not self.training and not torch.is_grad_enabled():
            folded = self.mlp.folded()
            if _mlp.vote_fast_ok(seed_features, folded):      # MLP + (vote = seed + y) in one launch
                return _mlp.vote_mlp_fast(seed_xyz, seed_features, folded)
            y = _mlp.pointwise_mlp(seed_features, folded, last_relu=False, want_cl=False)
        else:
            y = self.mlp(seed_features.unsqueeze(-1)).squeeze(-1)
        vote_xyz = (seed_xyz + y[:, :3, :].transpose(1, 2)).contiguous()
        vote_features = (seed_features + y[:, 3:, :]).contiguous()
        return vote_xyz, vote_features


class SizeAdaptiveAggregation(nn.Module):
    """3DSAD vote aggregation (a7): FPS over votes -> cluster centres -> per-cluster radius from
    the predicted object size -> adaptive ball query -> group -> MLP + max-pool."""

    def __init__(self, npoint: int = 256, nsample: int = 16, seed_feature_dim: int = 256,
                 mlp: Sequence[int] = (128, 128, 128), alpha: float = 1.0, r_min: float = 0.1,
                 r_max: float = 1.2, bn: bool = True, size_scale: float = 1.0, size_clip: float = 2.0):
        super().__init__()
        self.alpha, self.r_min, self.r_max = alpha, r_min, r_max
        self.size_scale, self.size_clip = size_scale, size_clip
        self.sa = PointnetSAModuleVotes(npoint, None, nsample, [seed_feature_dim] + list(mlp), bn=bn)
        # size head: the size that drives the per-cluster radius, predicted from the vote features
        # at the cluster centres: 2-layer 1x1 MLP -> log-size, size = size_scale * exp(clip(y)) [formula unpinned]
        self.size_mlp = SharedMLP([seed_feature_dim, 128, 3], bn=bn, last_relu=False)
        self.size_mlp.bns[-1] = nn.Identity()
        self.size_mlp.convs[-1] = nn.Conv2d(128, 3, kernel_size=1, bias=True)

    def predict_size(self, vote_features, cinds):
        """vote_features (B,C,n), cluster centre indices (B,K) -> predicted box size (B,K,3)."""
        centre = ops.gather_operation(vote_features.contiguous(), cinds)               # (B,C,K)
        if not self.training and not torch.is_grad_enabled():
            y = _mlp.pointwise_mlp(centre, self.size_mlp.folded(), last_relu=False, want_cl=False)
        else:
            y = self.size_mlp(centre.unsqueeze(-1)).squeeze(-1)
        y = y.clamp(min=-self.size_clip, max=self.size_clip)
        return (self.size_scale * torch.exp(y)).transpose(1, 2).contiguous()

    def forward(self, vote_xyz, vote_features, size=None):
        """`size` (B,K,3): externally predicted box sizes, or None: the module's own size head."""
        cinds = None
        if size is None:
            cinds = ops.furthest_point_sample(vote_xyz, self.sa.npoint)
            size = self.predict_size(vote_features, cinds)
        radius_t = ops.size_to_radius(size, self.alpha, self.r_min, self.r_max)
        cxyz, cfeat, cinds = self.sa(vote_xyz, vote_features, inds=cinds, radius_t=radius_t)
        return cxyz, cfeat, cinds, radius_t, size


class SADHotPath(nn.Module):
    """The timed unit 'scene' (SURVEY call stack 3): backbone -> voting -> size-adaptive
    vote aggregation.  `size` (B,256,3) is the predicted box size per cluster: passed in (the benchmark's synthetic
    sizes, a downstream proposal head) or, when None, predicted by the aggregation module's own size head."""

    def __init__(self, input_feature_dim: int = 1, bn: bool = True, mlp_dtype: str = "bf16"):
        super().__init__()
        self.backbone = Pointnet2Backbone(input_feature_dim, bn=bn)
        self.vgen = VotingModule(256, bn=bn)
        c = LAYER_CFG
        self.agg = SizeAdaptiveAggregation(c["agg"][0], c["agg"][2], 256, alpha=c["alpha"],
                                           r_min=c["r_min"], r_max=c["r_max"], bn=bn,
                                           size_scale=c["size_scale"], size_clip=c["size_clip"])
        self.set_mlp_dtype(mlp_dtype)

    def set_mlp_dtype(self, mlp_dtype: str):
        """Operand precision of every fused MLP stage: "bf16" (default; 2e-2 bar) or "tf32" (fp32 activations)."""
        if mlp_dtype not in ("bf16", "tf32"):
            raise ValueError(f"mlp_dtype must be 'bf16' or 'tf32', got {mlp_dtype!r}")
        for mod in self.modules():
            if isinstance(mod, SharedMLP):
                mod.mlp_dtype = mlp_dtype
        self.mlp_dtype = mlp_dtype
        return self

    @torch.no_grad()
    def load_params(self, params):
        """params: dict from config.make_params (BN folded)."""
        bb = self.backbone
        for name in ("sa1", "sa2", "sa3", "sa4"):
            getattr(bb, name).mlp_module.load_folded(params[name])
        bb.fp1.mlp.load_folded(params["fp1"])
        bb.fp2.mlp.load_folded(params["fp2"])
        self.vgen.mlp.load_folded(params["vote"])
        self.agg.sa.mlp_module.load_folded(params["agg"])
        if "size" in params:
            self.agg.size_mlp.load_folded(params["size"])
        return self

    def forward(self, xyz, features, size=None, ready_event=None):
        """`ready_event` (optional): a CUDA event after which the inputs are valid.  Passing it lets
        the coordinate-only chain of this batch start while the previous batch is still in its
        feature stages (two batches in flight); without it the call orders against the stream."""
        end = self.backbone(xyz, features, ready_event=ready_event)
        vxyz, vfeat = self.vgen(end["fp2_xyz"], end["fp2_features"])
        cxyz, cfeat, cinds, radius_t, size = self.agg(vxyz, vfeat, size)
        end.update(vote_xyz=vxyz, vote_features=vfeat, cluster_xyz=cxyz, cluster_features=cfeat,
                   cluster_inds=cinds, cluster_radius=radius_t, cluster_size=size)
        return end

    @torch.no_grad()
    def forward_host_async(self, xyz_host, feat_host, size_host, out_host):
        """End-to-end call with HOST (pinned) buffers, asynchronous: the H2D copies run on a copy
        stream (so they overlap the previous batch), the forward is ordered after them by an event,
        and the D2H of (cluster_xyz, cluster_features) into the pinned `out_host` pair is queued
        behind it.  Returns (out_host, done_event); `out_host` is valid after done_event."""
        dev = next(self.parameters()).device
        main = torch.cuda.current_stream(dev)
        copy = getattr(self, "_copy_stream", None)
        if copy is None or copy.device != dev:
            copy = self._copy_stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(copy):
            xyz = xyz_host.to(dev, non_blocking=True)
            feat = feat_host.to(dev, non_blocking=True)
            size = size_host.to(dev, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(copy)
        main.wait_event(ready)
        for t in (xyz, feat, size):
            t.record_stream(main)
        end = self.forward(xyz, feat, size, ready_event=ready)
        out_host[0].copy_(end["cluster_xyz"], non_blocking=True)
        out_host[1].copy_(end["cluster_features"], non_blocking=True)
        done = torch.cuda.Event()
        done.record(main)
        return out_host, done

    def make_host_outputs(self, batch: int):
        npoint = self.agg.sa.npoint
        c = self.agg.sa.mlp_module.channels[-1]
        return (torch.empty((batch, npoint, 3), dtype=torch.float32, pin_memory=True),
                torch.empty((batch, c, npoint), dtype=torch.float32, pin_memory=True))

    @torch.no_grad()
    def forward_host(self, xyz_host, feat_host, size_host, out_host=None):
        """Synchronous form of forward_host_async.  Returns (cluster_xyz_host, cluster_features_host)."""
        if out_host is None:
            out_host = self.make_host_outputs(xyz_host.shape[0])
        out_host, done = self.forward_host_async(xyz_host, feat_host, size_host, out_host)
        done.synchronize()
        return out_host
