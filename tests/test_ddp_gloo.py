"""world_size-2 checks of the data-parallel host logic on CPU (gloo): the flat-gradient all-reduce + 1/world scaling
equals DistributedDataParallel's gradient mean, the logging reduce matches `lib/distributed.py:260-270`, per-rank data
sharding is disjoint, and the reference arm of bench.py runs on rank 0 only."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from pointcontrast_b200 import trainer as T
    assert T.get_world_size() == world and T.get_rank() == rank
    # gradient mean: sum all-reduce of the flat buffer, then the SGD kernel's grad_scale = 1/world
    g = torch.arange(10, dtype=torch.float32) * (rank + 1)
    flat = g.clone()
    dist.all_reduce(flat)
    mean = flat * (1.0 / world)
    expect = torch.arange(10, dtype=torch.float32) * sum(r + 1 for r in range(world)) / world
    ok1 = torch.allclose(mean, expect)
    # logging reduce
    res = T.scaled_all_reduce_dict({"loss": torch.tensor(float(rank + 1)), "pos_loss": torch.tensor(2.0 * (rank + 1))}, world)
    ok2 = abs(float(res["loss"]) - 1.5) < 1e-6 and abs(float(res["pos_loss"]) - 3.0) < 1e-6
    q.put((rank, ok1, ok2))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_mean_and_logging_reduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 500
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    out = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(60)
    assert sorted(o[0] for o in out) == [0, 1] and all(o[1] and o[2] for o in out)


def test_rank_shards_are_disjoint_scene_pairs():
    from pointcontrast_b200 import synth
    a = synth.synth_pair(1000 * (100 * 0 + 0) + 0, scale=0.1)
    b = synth.synth_pair(1000 * (100 * 1 + 0) + 0, scale=0.1)
    assert a["coords0"].shape != b["coords0"].shape or (a["coords0"] != b["coords0"]).any()


def test_reference_arm_runs_on_rank0_only():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0", "--workload", "c0"], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip() == ""
