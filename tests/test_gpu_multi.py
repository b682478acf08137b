"""On-hardware data-parallel equivalence (SURVEY.md section 4 iii): 3 optimizer steps on N GPUs (one process per GPU, NCCL,
batch sharded over the ranks) leave every rank with the SAME parameters, equal to a single-GPU run over the whole batch.
Skipped when fewer than 2 GPUs are visible."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


FUSED_CASE = dict(latent=(8, 8), n=300, d=2, B=2, cin=2, cout=1, seed=1,
                  magno=dict(coord_dim=2, radius=0.35, hidden_size=64, mlp_layers=2, lifting_channels=16, neighbor_search_method="native"),
                  transformer=dict(patch_size=2, hidden_size=64, num_layers=1, ffn_multiplier=2), attn=dict(num_heads=2, num_kv_heads=2))


def _shard_worker(rank, world, port, out):
    """bf16 mode with the fused edge-weight kernels: 2 steps with and without row sharding must give the same parameters."""
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import cases
    from tests.helpers import build_model
    from gaot_b200 import ops
    from gaot_b200.trainer import TrainStep
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ops.set_precision("bf16")
    ops.SHARD_EDGES_MIN = 128
    c = FUSED_CASE
    lat, x = cases.latent_grid(c["latent"]), cases.uniform_points(c["n"], 2, 5)
    pn, tgt = cases.normal((2 * world, c["n"], c["cin"]), 77), cases.normal((2 * world, c["n"], c["cout"]), 78)
    res = {}
    for shard in (False, True):
        ops.enable_edge_sharding(True if shard else None)
        torch.manual_seed(0)
        model = build_model(c, dev)
        step = TrainStep(model, lr=1e-3, weight_decay=1e-2)
        calls = []
        from gaot_b200 import _lib
        for _ in range(2):
            step.fixed_coords_step(lat.to(dev), x.to(dev), pn[rank * 2:(rank + 1) * 2].to(dev), tgt[rank * 2:(rank + 1) * 2].to(dev))
        res[shard] = step.flat_param.detach().clone()
    ops.enable_edge_sharding(None)
    gathered = [torch.empty_like(res[True]) for _ in range(world)]
    dist.all_gather(gathered, res[True])
    if rank == 0:
        out["ranks_equal"] = all(torch.equal(gathered[0], t) for t in gathered[1:])
        out["rel"] = float((res[True] - res[False]).double().norm() / res[False].double().norm())
        out["moved"] = float((res[False] - gathered[0]).abs().max())
    dist.barrier()
    dist.destroy_process_group()


def _worker(rank, world, port, precision, out):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import cases
    from tests.helpers import build_model, golden_params, load_golden
    from gaot_b200 import ops
    from gaot_b200.trainer import TrainStep
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    ops.set_precision(precision)
    name = "tiny2d"
    c = dict(cases.CASES[name])
    g = load_golden(name)
    model = build_model(c, dev)
    model.load_state_dict({k: v.to(dev) for k, v in golden_params(g).items()})
    lat, x, _, _ = cases.case_inputs(name)
    Bg = 2 * world                                         # global batch, sharded contiguously
    pn = cases.normal((Bg, c["n"], c["cin"]), 77)
    tgt = cases.normal((Bg, c["n"], c["cout"]), 78)
    per = Bg // world
    step = TrainStep(model, lr=1e-3, weight_decay=1e-2)
    for _ in range(3):
        step.fixed_coords_step(lat.to(dev), x.to(dev), pn[rank * per:(rank + 1) * per].to(dev), tgt[rank * per:(rank + 1) * per].to(dev))
    flat = step.flat_param.detach().clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    if rank == 0:
        out["equal"] = all(torch.equal(gathered[0], t) for t in gathered[1:])
        out["flat"] = flat.cpu()
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_data_parallel_equals_single_gpu(precision):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    from oracle import cases
    from tests.helpers import build_model, golden_params, load_golden, rel_l2
    from gaot_b200 import ops
    from gaot_b200.trainer import TrainStep
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), precision, out), nprocs=world, join=True)
    assert out["equal"], "ranks diverged: parameters differ between GPUs after 3 steps"
    # the same 3 steps on ONE GPU over the whole batch (mean-of-means is exact for equal shards)
    ops.set_precision(precision)
    try:
        name = "tiny2d"
        c = cases.CASES[name]
        g = load_golden(name)
        model = build_model(c)
        model.load_state_dict({k: v.cuda() for k, v in golden_params(g).items()})
        lat, x, _, _ = cases.case_inputs(name)
        pn, tgt = cases.normal((2 * world, c["n"], c["cin"]), 77), cases.normal((2 * world, c["n"], c["cout"]), 78)
        step = TrainStep(model, lr=1e-3, weight_decay=1e-2, distributed=False)
        for _ in range(3):
            step.fixed_coords_step(lat.cuda(), x.cuda(), pn.cuda(), tgt.cuda())
    finally:
        ops.set_precision("fp32")
    tol = 2e-5 if precision == "fp32" else 2e-3        # bf16: the two runs round different partial sums
    assert rel_l2(out["flat"], step.flat_param) < tol


def test_row_sharded_magno_equals_replicated():
    """MAGNO rows sharded over the ranks (every rank evaluates the edge-weight MLP, the gather / reduce and their backward for
    its half of the latent / query rows and ALL samples; all-gather in, all-to-all out over NCCL) trains to the same
    parameters as plain data parallelism, on every rank."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_shard_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out["ranks_equal"]
    assert out["rel"] < 2e-4, out["rel"]              # bf16 mode: the two runs round different partial sums
