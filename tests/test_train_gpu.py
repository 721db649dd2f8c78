"""BASELINE config 4 (training step): forward + backward of the whole hot path through the torch
composition and this library's scatter-add backward kernels, an optimizer step, and the same step under
DistributedDataParallel (the only collective of the path: the NCCL gradient all-reduce)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _batch(B=2, N=3000):
    from sad_b200.config import LAYER_CFG
    from sad_b200.scenes import make_scenes, make_sizes
    xyz, feat = make_scenes(B, N, "surface")
    size = make_sizes(B, LAYER_CFG["agg"][0])
    return tuple(torch.from_numpy(a).to(DEV) for a in (xyz, feat, size))


def _loss(end):
    # synthetic scalar loss touching both heads: cluster features and the vote offsets
    return end["cluster_features"].square().mean() + 0.1 * (end["vote_xyz"] - end["fp2_xyz"]).square().mean()


def test_train_step_backprops_through_every_stage_and_learns():
    import sad_b200  # noqa: F401
    from sad_b200.config import make_params
    from sad_b200.modules import SADHotPath
    torch.manual_seed(0)
    model = SADHotPath(1).load_params(make_params(0)).to(DEV).train()
    opt = torch.optim.SGD(model.parameters(), lr=1e-2)
    xyz, feat, size = _batch()
    losses = []
    for step in range(4):
        opt.zero_grad(set_to_none=True)
        end = model(xyz, feat, size)
        loss = _loss(end)
        loss.backward()
        if step == 0:
            for name, p in model.named_parameters():
                assert p.grad is not None and torch.isfinite(p.grad).all(), name
            convs = [p for n, p in model.named_parameters() if n.endswith("convs.0.weight")]
            assert all(float(p.grad.abs().sum()) > 0 for p in convs)       # gradient reaches the first layer of every stage
        opt.step()
        losses.append(float(loss.detach()))
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    # sampling is coordinate-only: training does not change the indices
    model.eval()
    with torch.no_grad():
        assert torch.equal(model(xyz, feat, size)["sa1_inds"], end["sa1_inds"])


def test_ddp_wrapped_step_matches_plain_step():
    """world_size 1 NCCL group: exercises the DDP gradient bucket / all-reduce hooks on the real backend."""
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    import sad_b200  # noqa: F401
    from sad_b200.config import make_params
    from sad_b200.modules import SADHotPath
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device(DEV))
    try:
        xyz, feat, size = _batch(B=1, N=2500)
        grads = []
        for wrap in (False, True):
            torch.manual_seed(0)
            model = SADHotPath(1).load_params(make_params(0)).to(DEV).train()
            net = DDP(model, device_ids=[0]) if wrap else model
            _loss(net(xyz, feat, size)).backward()
            grads.append(torch.cat([p.grad.flatten() for p in model.parameters()]))
        # same math; scatter-add atomics reorder fp32 sums (and train-mode BN amplifies the noise), hence a norm-wise bar
        rel = float((grads[0] - grads[1]).norm() / grads[0].norm())
        assert rel < 1e-2, rel
        assert grads[0].numel() > 900_000        # ~3.8 MB of fp32 gradients: the one collective of the path
    finally:
        dist.destroy_process_group()


def test_train_step_under_bf16_autocast():
    """Config 4 asks for a bf16 forward+backward: under autocast the point-wise MLPs run in bf16 while the
    differentiable library ops take their inputs back to fp32 (ops._amp_fwd) and return fp32 gradients."""
    import sad_b200  # noqa: F401
    from sad_b200.config import make_params
    from sad_b200.modules import SADHotPath
    torch.manual_seed(0)
    model = SADHotPath(1).load_params(make_params(0)).to(DEV).train()
    xyz, feat, size = _batch(B=2, N=2500)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        end = model(xyz, feat, size)
    _loss({k: (v.float() if torch.is_floating_point(v) else v) for k, v in end.items() if torch.is_tensor(v)}).backward()
    for name, p in model.named_parameters():
        assert p.grad is not None and p.grad.dtype == torch.float32 and torch.isfinite(p.grad).all(), name
    # indices do not depend on the precision of the features
    model.eval()
    with torch.no_grad():
        assert torch.equal(model(xyz, feat, size)["sa1_inds"], end["sa1_inds"])
