"""Pipelined executor for the hot path: CUDA graphs + streams, several batches in flight.

The hot path is a chain of ~60 short kernels per batch whose critical path is the serial
furthest-point-sampling chain (hard part H3): run one
batch at a time and most SMs idle behind that chain.  So the executor

  * captures the whole forward of one batch (all kernels of libsad_b200 plus the torch glue,
    including the side-stream fork/join of the coordinate-only chain) into ONE CUDA graph
    per slot, with static input / output buffers -- no Python, allocator or launch overhead
    at run time;
  * owns `slots` such graphs, each on its own stream, and feeds them round-robin, so the
    FPS chain of batch k+1 overlaps the ball queries / tensor-core MLPs of batch k;
  * for host callers stages the inputs through pinned buffers: H2D copy, graph launch and
    D2H copy of the result are queued back to back on the slot's stream.

Results are delivered in submission order and are bit-identical to SADHotPath.forward
(the graph replays exactly the kernels the eager call launches).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import _lib


class _Slot:
    __slots__ = ("stream", "xyz", "feat", "size", "graph", "end", "out_host", "done", "busy", "launches")


class PipelinedHotPath:
    """`slots` CUDA-graph instances of `model.forward` for a fixed (batch, n_points) shape.

        eng = PipelinedHotPath(model, batch=8, n_points=40000)
        t = eng.submit_host(xyz_h, feat_h, size_h)      # pinned host tensors; returns a ticket
        cluster_xyz_h, cluster_feat_h = eng.result(t)   # pinned host views, valid until the slot is reused
    """

    def __init__(self, model, batch: int, n_points: int, feat_dim: int = 1, slots: int = 3,
                 device: Optional[torch.device] = None, warmup: int = 2, fps_policy: str = "throughput",
                 mlp_tiles_per_cta: int = 6):
        """`fps_policy`: "throughput" (default) captures the one-SM-per-scene FPS kernel -- longer per batch, but
        about half the SM-time, which is what bounds a pipeline with several batches in flight; "latency" captures the
        cluster kernel (shortest time per batch; right for 1-2 slots)."""
        self.model = model
        self.fps_policy = fps_policy
        self.mlp_tiles_per_cta = mlp_tiles_per_cta      # narrower fused-MLP grids for the small stages (less SM-time)
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        self.batch, self.n_points, self.feat_dim = batch, n_points, feat_dim
        self.n_clusters = model.agg.sa.npoint
        self.launches_per_batch = 0
        self._next = 0
        self._slots: List[_Slot] = []
        lib = _lib.load()
        dev = self.device
        from . import modules as _modules, mlp as _mlp
        saved_policy = _modules.FPS_POLICY[0]
        _modules.FPS_POLICY[0] = fps_policy
        lib.sad_mlp_set_tiles_per_cta(int(mlp_tiles_per_cta))
        _mlp.TILES_PER_CTA[0] = int(mlp_tiles_per_cta)
        try:
            self._capture(model, batch, n_points, feat_dim, slots, dev, warmup, lib)
        finally:
            _modules.FPS_POLICY[0] = saved_policy
            lib.sad_mlp_set_tiles_per_cta(1)
            _mlp.TILES_PER_CTA[0] = 1
        self.launches_per_batch = self._slots[0].launches

    def _capture(self, model, batch, n_points, feat_dim, slots, dev, warmup, lib):
        with torch.cuda.device(dev), torch.no_grad():
            for _ in range(slots):
                s = _Slot()
                s.stream = torch.cuda.Stream(device=dev)
                s.xyz = torch.zeros((batch, n_points, 3), dtype=torch.float32, device=dev)
                s.feat = torch.zeros((batch, feat_dim, n_points), dtype=torch.float32, device=dev)
                s.size = torch.ones((batch, self.n_clusters, 3), dtype=torch.float32, device=dev)
                s.busy = False
                s.done = torch.cuda.Event()
                self._slots.append(s)
            # well-formed warm-up input (weights get packed, kernels configured) before capture
            g = torch.Generator(device="cpu").manual_seed(0)
            wx = (torch.rand((batch, n_points, 3), generator=g) * 4).to(dev)
            for s in self._slots:
                s.xyz.copy_(wx)
                s.feat.copy_(wx[:, :, 2].unsqueeze(1).expand(-1, feat_dim, -1))
            torch.cuda.synchronize(dev)
            for s in self._slots:
                with torch.cuda.stream(s.stream):
                    for _ in range(max(1, warmup)):
                        model(s.xyz, s.feat, s.size)
                s.stream.synchronize()
                s.graph = torch.cuda.CUDAGraph()
                l0 = lib.sad_launch_count()
                with torch.cuda.graph(s.graph, stream=s.stream):
                    s.end = model(s.xyz, s.feat, s.size)
                s.launches = int(lib.sad_launch_count() - l0)
                s.out_host = model.make_host_outputs(batch)
            torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------ submission
    @property
    def slots(self) -> int:
        return len(self._slots)

    def _acquire(self) -> Tuple[int, _Slot]:
        i = self._next
        self._next = (i + 1) % len(self._slots)
        s = self._slots[i]
        if s.busy:
            s.done.synchronize()      # back-pressure: the slot's previous batch must have left the GPU
        s.busy = True
        return i, s

    @torch.no_grad()
    def submit_device(self, xyz, feat, size, after: Optional[torch.cuda.Event] = None, to_host: bool = False) -> int:
        """Inputs already resident in HBM: device-to-device copy into the slot, then the graph.  The slot's stream is
        ordered after the caller's current stream (or after `after`, an event the caller recorded once the inputs were
        valid), and the inputs are kept alive for the copy (record_stream), so tensors produced just before the call
        or dropped right after it are safe."""
        i, s = self._acquire()
        if after is None:
            after = torch.cuda.Event()
            after.record(torch.cuda.current_stream(self.device))
        for t in (xyz, feat, size):
            if t.is_cuda:
                t.record_stream(s.stream)
        with torch.cuda.stream(s.stream):
            s.stream.wait_event(after)
            s.xyz.copy_(xyz, non_blocking=True)
            s.feat.copy_(feat, non_blocking=True)
            s.size.copy_(size, non_blocking=True)
            s.graph.replay()
            if to_host:
                s.out_host[0].copy_(s.end["cluster_xyz"], non_blocking=True)
                s.out_host[1].copy_(s.end["cluster_features"], non_blocking=True)
            s.done.record(s.stream)
        return i

    @torch.no_grad()
    def submit_host(self, xyz_host, feat_host, size_host, after: Optional[torch.cuda.Event] = None) -> int:
        """Host (pinned) buffers in, pinned host results out; everything queued on the slot's stream."""
        i, s = self._acquire()
        with torch.cuda.stream(s.stream):
            if after is not None:
                s.stream.wait_event(after)
            s.xyz.copy_(xyz_host, non_blocking=True)
            s.feat.copy_(feat_host, non_blocking=True)
            s.size.copy_(size_host, non_blocking=True)
            s.graph.replay()
            s.out_host[0].copy_(s.end["cluster_xyz"], non_blocking=True)
            s.out_host[1].copy_(s.end["cluster_features"], non_blocking=True)
            s.done.record(s.stream)
        return i

    def result(self, ticket: int):
        """Block until batch `ticket` is done; returns the slot's pinned (cluster_xyz, cluster_features)."""
        s = self._slots[ticket]
        s.done.synchronize()
        s.busy = False
        return s.out_host

    def outputs(self, ticket: int) -> dict:
        """Device-side end-points dict of the slot (static buffers, valid until the slot is reused)."""
        return self._slots[ticket].end

    def join(self, stream: Optional[torch.cuda.Stream] = None):
        """Make `stream` (default: current) wait for every batch submitted so far."""
        stream = stream or torch.cuda.current_stream(self.device)
        for s in self._slots:
            if s.busy:
                stream.wait_event(s.done)

    def drain(self):
        for s in self._slots:
            if s.busy:
                s.done.synchronize()
                s.busy = False
