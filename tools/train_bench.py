"""BASELINE config 4: full forward + backward + optimizer step of the hot path, data-parallel with the
NCCL gradient all-reduce (the path's only collective).  One process per GPU:

    python tools/train_bench.py [--B 8] [--N 20000] [--steps 10] [--warmup 3] [--amp]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/train_bench.py

Forward goes through the torch composition of the modules (train mode: batch-norm statistics, autograd), the
sampling / search / grouping / interpolation ops and their scatter-add backward kernels through libsad_b200.
Prints one JSON line on rank 0 (device-timed, max over ranks)."""
import argparse, json, os, sys
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sad_b200  # noqa: E402,F401
from sad_b200.config import LAYER_CFG, make_params  # noqa: E402
from sad_b200.modules import SADHotPath  # noqa: E402
from sad_b200.scenes import make_scenes, make_sizes  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--N", type=int, default=20000)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--amp", action="store_true", help="bf16 autocast around the point-wise MLPs")
    ap.add_argument("--deterministic", action="store_true",
                    help="sort-by-destination scatter-add backward (csrc/scatter.cu) instead of fp32 atomics")
    a = ap.parse_args()
    from sad_b200 import dist as D, ops as _ops
    _ops.set_deterministic(a.deterministic)
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = SADHotPath(1).load_params(make_params(0)).to(dev).train()
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=True) if world > 1 else model
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9)
    first = D.weak_scaling_first_scene(rank, a.B)
    xyz, feat = make_scenes(a.B, a.N, "surface", first_scene=first)
    size = make_sizes(a.B, LAYER_CFG["agg"][0], first_scene=first)
    x, f, s = (torch.from_numpy(t).to(dev) for t in (xyz, feat, size))
    n_param = sum(p.numel() for p in model.parameters())

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=a.amp):
            end = net(x, f, s)
        loss = end["cluster_features"].float().square().mean() + \
            0.1 * (end["vote_xyz"] - end["fp2_xyz"]).float().square().mean()
        loss.backward()
        opt.step()
        return loss

    for _ in range(a.warmup):
        loss = step()
    torch.cuda.synchronize()
    D.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    (ms,) = D.reduce_scalars([e0.elapsed_time(e1)], "max", device=dev)
    if rank == 0:
        per = float(ms) / a.steps
        print(json.dumps({"config": "configs[3]: forward+backward+SGD step, DDP (NCCL gradient all-reduce)",
                          "n_gpus": world, "scenes_per_gpu_per_step": a.B, "points_per_scene": a.N,
                          "ms_per_step": round(per, 3), "train_scenes_per_s": round(world * a.B / per * 1e3, 1),
                          "mlp_dtype": "bf16 autocast" if a.amp else "f32 (TF32 off)",
                          "allreduce_bytes_per_step": 4 * n_param if world > 1 else 0,
                          "scatter_add_backward": "deterministic (sorted segments)" if a.deterministic else "fp32 atomics",
                          "final_loss": float(loss.detach())}), flush=True)
    D.barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
