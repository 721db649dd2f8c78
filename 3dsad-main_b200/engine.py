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
    __slots__ = ("stream", "xyz", "feat", "size", "graph", "end", "out_host", "done", "busy", "launches", "desc", "ready", "hold")


class PipelinedHotPath:
    """`slots` CUDA-graph instances of `model.forward` for a fixed (batch, n_points) shape.

        eng = PipelinedHotPath(model, batch=8, n_points=40000)
        t = eng.submit_host(xyz_h, feat_h, size_h)      # pinned host tensors; returns a ticket
        cluster_xyz_h, cluster_feat_h = eng.result(t)   # pinned host views, valid until the slot is reused
    """

    def __init__(self, model, batch: int, n_points: int, feat_dim: int = 1, slots: int = 3,
                 device: Optional[torch.device] = None, warmup: int = 2, fps_policy: str = "throughput",
                 mlp_tiles_per_cta: int = 6, native_submit: bool = True, linear_graph: bool = True):
        """`linear_graph`: capture the batch on ONE stream (no side-stream fork / join of the coordinate-only chain), so
        the graph is a straight line of kernel nodes: the driver launches those in near-constant time, a forked graph
        costs ~100 us of host time per launch -- more than the GPU needs to start the batch.  With several batches in
        flight the intra-batch overlap the fork buys is provided by the other batches anyway; for 1-2 slots pass False
        (shorter single-batch latency).
        `native_submit`: queue a batch (event wait, input copies, graph launch, result copies, completion event)
        with ONE call into libsad_b200 (`sad_engine_submit`) instead of ~10 PyTorch calls; inputs whose shape / dtype /
        layout differ from the slot's static buffers take the PyTorch path (copy_ converts).
        `fps_policy`: "throughput" (default) captures the one-SM-per-scene FPS kernel -- longer per batch, but
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
        _mlp.TILES_PER_CTA[0] = int(mlp_tiles_per_cta)
        self.linear_graph = bool(linear_graph)
        backbone = getattr(model, "backbone", None)
        saved_overlap = getattr(backbone, "overlap_geometry", None)
        if self.linear_graph and saved_overlap is not None:
            backbone.overlap_geometry = False
        try:
            self._capture(model, batch, n_points, feat_dim, slots, dev, warmup, lib)
        finally:
            _modules.FPS_POLICY[0] = saved_policy
            _mlp.TILES_PER_CTA[0] = 1
            if saved_overlap is not None:
                backbone.overlap_geometry = saved_overlap
        self.launches_per_batch = self._slots[0].launches
        self._lib = lib
        self.native_submit = bool(native_submit)
        if self.native_submit:
            self._prepare_native()

    def _prepare_native(self):
        """Per slot: the constant part of its sad_submit_desc (graph, destinations, sizes, result copies, completion
        event).  torch creates an Event's cudaEvent_t on its first record, so every event is recorded once here."""
        import ctypes
        with torch.cuda.device(self.device):
            for s in self._slots:
                s.done.record(s.stream)
                s.ready = torch.cuda.Event()
                s.ready.record(s.stream)
                d = _lib.SubmitDesc()
                try:
                    d.graph_exec = int(s.graph.raw_cuda_graph_exec())
                except RuntimeError:                      # a keep_graph=True capture is instantiated on demand
                    s.graph.instantiate()
                    d.graph_exec = int(s.graph.raw_cuda_graph_exec())
                d.done_event = int(s.done.cuda_event)
                d.n_in = 3
                for i, t in enumerate((s.xyz, s.feat, s.size)):
                    d.in_dst[i] = t.data_ptr()
                    d.in_bytes[i] = t.numel() * t.element_size()
                outs = (s.end["cluster_xyz"], s.end["cluster_features"])
                for i, (h, t) in enumerate(zip(s.out_host, outs)):
                    if not (t.is_contiguous() and h.is_contiguous() and h.shape == t.shape and h.dtype == t.dtype):
                        self.native_submit = False       # results need a converting copy: PyTorch path
                        return
                    d.out_dst[i] = h.data_ptr()
                    d.out_src[i] = t.data_ptr()
                    d.out_bytes[i] = t.numel() * t.element_size()
                s.desc = d
            torch.cuda.synchronize(self.device)
        self._byref = ctypes.byref

    def _native_ok(self, s, xyz, feat, size) -> bool:
        return (self.native_submit and xyz.shape == s.xyz.shape and feat.shape == s.feat.shape and size.shape == s.size.shape
                and xyz.dtype == feat.dtype == size.dtype == torch.float32
                and xyz.is_contiguous() and feat.is_contiguous() and size.is_contiguous())

    def _submit_native(self, s, xyz, feat, size, wait_event, to_host: bool):
        # the copies are raw cudaMemcpyAsync calls: nothing tells PyTorch's pinned-host / device allocators that the
        # sources are in use, so the slot keeps them alive until its batch has left the GPU (result / reuse / drain)
        s.hold = (xyz, feat, size)
        d = s.desc
        d.in_src[0], d.in_src[1], d.in_src[2] = xyz.data_ptr(), feat.data_ptr(), size.data_ptr()
        d.wait_event = int(wait_event.cuda_event) if wait_event is not None else None
        d.n_out = 2 if to_host else 0
        _lib.check(self._lib.sad_engine_submit(self._byref(d), s.stream.cuda_stream), "sad_engine_submit")

    def _capture(self, model, batch, n_points, feat_dim, slots, dev, warmup, lib):
        with torch.cuda.device(dev), torch.no_grad():
            for _ in range(slots):
                s = _Slot()
                s.stream = torch.cuda.Stream(device=dev)
                s.xyz = torch.zeros((batch, n_points, 3), dtype=torch.float32, device=dev)
                s.feat = torch.zeros((batch, feat_dim, n_points), dtype=torch.float32, device=dev)
                s.size = torch.ones((batch, self.n_clusters, 3), dtype=torch.float32, device=dev)
                s.busy = False
                s.hold = None
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
            # the first launch of an instantiated graph uploads it to the device (~60 us of host time per graph,
            # measured): pay that here, once per slot, not inside the caller's first `slots` submissions
            for s in self._slots:
                with torch.cuda.stream(s.stream):
                    s.graph.replay()
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
        s.hold = None
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
            after = s.ready if self.native_submit else torch.cuda.Event()
            after.record(torch.cuda.current_stream(self.device))
        for t in (xyz, feat, size):
            if t.is_cuda:
                t.record_stream(s.stream)
        if self._native_ok(s, xyz, feat, size):
            self._submit_native(s, xyz, feat, size, after, to_host)
            return i
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

    def slot_inputs(self, index: int):
        """The static device input buffers (xyz, feat, size) of slot `index`: a producer that can write a batch
        straight into them (a decoder, a previous pipeline stage) skips the copy of submit_device; `next_slot` is the
        slot the next submit_* call takes."""
        s = self._slots[index]
        return s.xyz, s.feat, s.size

    @property
    def next_slot(self) -> int:
        return self._next

    @torch.no_grad()
    def submit_resident(self, after: Optional[torch.cuda.Event] = None, to_host: bool = False) -> int:
        """Zero-copy submission: run the next slot's graph on what its input buffers (slot_inputs) hold.  The buffers
        must have been written on a stream the slot's stream is ordered after: pass the event recorded after the
        writes as `after` (None: the caller guarantees they completed, e.g. they were filled before a synchronize)."""
        i, s = self._acquire()
        if self.native_submit:
            d = s.desc
            d.n_in = 0
            try:
                d.wait_event = int(after.cuda_event) if after is not None else None
                d.n_out = 2 if to_host else 0
                _lib.check(self._lib.sad_engine_submit(self._byref(d), s.stream.cuda_stream), "sad_engine_submit")
            finally:
                d.n_in = 3
            return i
        with torch.cuda.stream(s.stream):
            if after is not None:
                s.stream.wait_event(after)
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
        if self._native_ok(s, xyz_host, feat_host, size_host):
            self._submit_native(s, xyz_host, feat_host, size_host, after, True)
            return i
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
        s.hold = None
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
                s.hold = None


class ShardedHotPath:
    """Scene-data-parallel front end: ONE PROCESS PER GPU, every rank owns a contiguous block
    of the scenes and runs its own PipelinedHotPath over it; there is no data-path collective.  Launch with torchrun:

        torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 my_script.py
        ...
        dist.init_process_group("nccl")                      # or not at all: a single process owns every scene
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        shp = ShardedHotPath(model, batch=8, n_points=40000)
        cxyz, cfeat = shp.run(xyz_all, feat_all, size_all)   # host arrays of ALL scenes, the same on every rank
        # rank 0: (S,256,3) / (S,128,256) for all S scenes in scene order; other ranks: their own block

    `engine` may be any object with submit_host / result / slots (tests drive the sharding logic on CPU with a stub).
    Results are bit-identical to a single process running all scenes: every scene's result depends on that scene only.
    """

    def __init__(self, model=None, batch: int = 8, n_points: int = 40000, engine=None, **engine_kw):
        from . import dist as _dist
        self._dist = _dist
        self.world, self.rank = _dist.world_info()
        self.batch = batch
        self.engine = engine if engine is not None else PipelinedHotPath(model, batch, n_points, **engine_kw)

    def my_range(self, num_scenes: int):
        return self._dist.shard_range(num_scenes, self.world, self.rank)

    def run(self, xyz, feat, size, gather: bool = True):
        """xyz (S,N,3), feat (S,C,N), size (S,K,3): host tensors / arrays holding ALL scenes -> (cluster_xyz, cluster_features)
        host tensors: every scene on rank 0 (gather=True), this rank's block elsewhere / otherwise."""
        xyz, feat, size = (torch.as_tensor(a) for a in (xyz, feat, size))
        lo, hi = self.my_range(xyz.shape[0])
        outs_x, outs_f, tickets = [], [], []

        def drain_one():
            (k0, n_valid), t = tickets.pop(0)
            cx, cf = self.engine.result(t)
            outs_x.append(cx[:n_valid].clone())
            outs_f.append(cf[:n_valid].clone())

        for k0 in range(lo, hi, self.batch):
            k1 = min(hi, k0 + self.batch)
            sel = list(range(k0, k1)) + [k1 - 1] * (self.batch - (k1 - k0))      # pad the last batch with its last scene
            staged = tuple(a[sel].contiguous().pin_memory() if torch.cuda.is_available() else a[sel].contiguous()
                           for a in (xyz, feat, size))
            tickets.append(((k0, k1 - k0), self.engine.submit_host(*staged)))
            if len(tickets) >= self.engine.slots:
                drain_one()
        while tickets:
            drain_one()
        mine_x = torch.cat(outs_x) if outs_x else torch.empty((0,))
        mine_f = torch.cat(outs_f) if outs_f else torch.empty((0,))
        if not gather or self.world == 1:
            return mine_x, mine_f
        import torch.distributed as dist
        parts = [None] * self.world if self.rank == 0 else None
        dist.gather_object((mine_x, mine_f), parts, dst=0)       # control plane only: results leave the GPUs as host tensors
        if self.rank != 0:
            return mine_x, mine_f
        return (torch.cat([p[0] for p in parts if p[0].numel()]), torch.cat([p[1] for p in parts if p[1].numel()]))
