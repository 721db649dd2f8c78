"""What is in the per-batch CUDA graph, and what does queuing one batch cost on the host?
Captures the hot path like engine.PipelinedHotPath does (keep_graph=True so the cudaGraph_t can be inspected), lists
node types / edge structure, then times cudaGraphLaunch / cudaMemcpyAsync / event calls on an idle GPU.
    python tools/graph_probe.py [--forked]"""
import os, sys, time, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cuda.bindings import runtime as rt
import sad_b200  # noqa
from sad_b200.config import make_params
from sad_b200.engine import PipelinedHotPath
from sad_b200.modules import SADHotPath

forked = "--forked" in sys.argv
_Real = torch.cuda.CUDAGraph
torch.cuda.CUDAGraph = lambda *a, **k: _Real(keep_graph=True)
dev = torch.device("cuda:0")
model = SADHotPath(1).load_params(make_params(0)).to(dev).eval()
eng = PipelinedHotPath(model, 8, 40000, slots=24, device=dev, fps_policy="throughput", mlp_tiles_per_cta=6, linear_graph=not forked)
s = eng._slots[0]
g = int(s.graph.raw_cuda_graph())
err, _, n = rt.cudaGraphGetNodes(g, 0)
err, nodes, n = rt.cudaGraphGetNodes(g, n)
types = collections.Counter()
for nd in nodes:
    err, t = rt.cudaGraphNodeGetType(nd)
    types[str(t).split(".")[-1]] += 1
err, _, _, ne = rt.cudaGraphGetEdges(g, 0)
print(f"graph ({'forked' if forked else 'linear'}): {n} nodes {dict(types)}; {ne} edges (a straight line has nodes-1)", flush=True)

ex = int(s.graph.raw_cuda_graph_exec())
st = s.stream.cuda_stream


def timeit(name, fn, n=200):
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        if _ % 8 == 7:
            torch.cuda.synchronize()
    ts.sort()
    print(f"{name:40s} median {1e6 * ts[len(ts) // 2]:7.1f} us   p90 {1e6 * ts[int(len(ts) * .9)]:7.1f} us", flush=True)


x = torch.zeros_like(s.xyz)
timeit("cudaGraphLaunch (cuda-python)", lambda: rt.cudaGraphLaunch(ex, st))
timeit("graph.replay() (torch)", lambda: s.graph.replay())
timeit("cudaMemcpyAsync D2D 3.84 MB", lambda: rt.cudaMemcpyAsync(s.xyz.data_ptr(), x.data_ptr(), x.numel() * 4, rt.cudaMemcpyKind.cudaMemcpyDefault, st))
xh = torch.zeros(s.xyz.shape, pin_memory=True)
timeit("cudaMemcpyAsync H2D 3.84 MB pinned", lambda: rt.cudaMemcpyAsync(s.xyz.data_ptr(), xh.data_ptr(), x.numel() * 4, rt.cudaMemcpyKind.cudaMemcpyDefault, st))
ev = torch.cuda.Event(); ev.record()
timeit("cudaEventRecord", lambda: rt.cudaEventRecord(int(ev.cuda_event), st))
sx = (x, torch.zeros_like(s.feat), torch.zeros_like(s.size))
timeit("eng.submit_device (native)", lambda: (eng.submit_device(*sx), eng.drain() if eng._next == 0 else None))


# ---- the bench's fill: 20 batches queued back to back on a busy GPU, per-call host times
def fill(label, fn):
    eng.drain(); torch.cuda.synchronize()
    ts = []
    t_all = time.perf_counter()
    for k in range(20):
        t0 = time.perf_counter(); fn(k); ts.append(1e6 * (time.perf_counter() - t0))
    tot = 1e6 * (time.perf_counter() - t_all)
    eng.drain(); torch.cuda.synchronize()
    print(f"{label:34s} total {tot:7.0f} us; per call: " + " ".join(f"{t:.0f}" for t in ts), flush=True)


def raw(k, kind):
    sl = eng._slots[k]
    stq = sl.stream.cuda_stream
    for dst, src in zip((sl.xyz, sl.feat, sl.size), sx):
        rt.cudaMemcpyAsync(dst.data_ptr(), src.data_ptr(), src.numel() * 4, kind, stq)
    rt.cudaGraphLaunch(int(sl.graph.raw_cuda_graph_exec()), stq)
    rt.cudaEventRecord(int(sl.done.cuda_event), stq)


def graph_only(k):
    sl = eng._slots[k]
    rt.cudaGraphLaunch(int(sl.graph.raw_cuda_graph_exec()), sl.stream.cuda_stream)


for rep in range(2):
    fill("eng.submit_device", lambda k: eng.submit_device(*sx))
    fill("raw: 3 copies(Default)+graph+event", lambda k: raw(k, rt.cudaMemcpyKind.cudaMemcpyDefault))
    fill("raw: 3 copies(D2D)+graph+event", lambda k: raw(k, rt.cudaMemcpyKind.cudaMemcpyDeviceToDevice))
    fill("raw: graph launch only", graph_only)


# ---- real scenes resident in the slots (the bench's value loop): which call is slow?
from sad_b200.scenes import make_scenes, make_sizes
from sad_b200.config import LAYER_CFG
for i in range(eng.slots):
    xyz, feat = make_scenes(8, 40000, "surface", first_scene=8 * i)
    size = make_sizes(8, LAYER_CFG["agg"][0], first_scene=8 * i)
    for dst, src in zip(eng.slot_inputs(i), (xyz, feat, size)):
        dst.copy_(torch.from_numpy(src))
torch.cuda.synchronize()
ev0 = torch.cuda.Event(enable_timing=True)


def variant(k, wait, rec):
    sl = eng._slots[k]
    stq = sl.stream.cuda_stream
    if wait:
        rt.cudaStreamWaitEvent(stq, int(ev0.cuda_event), 0)
    rt.cudaGraphLaunch(int(sl.graph.raw_cuda_graph_exec()), stq)
    if rec:
        rt.cudaEventRecord(int(sl.done.cuda_event), stq)


for rep in range(2):
    ev0.record()
    fill("real: graph only", lambda k: variant(k, False, False))
    ev0.record()
    fill("real: wait + graph", lambda k: variant(k, True, False))
    ev0.record()
    fill("real: graph + record", lambda k: variant(k, False, True))
    ev0.record()
    fill("real: wait + graph + record", lambda k: variant(k, True, True))
    ev0.record()
    fill("real: eng.submit_resident(after)", lambda k: eng.submit_resident(after=ev0))
