"""World-size-2 (gloo, CPU) tests of bench.py's multi-rank bookkeeping: per-rank shards of the batch
of frame pairs, the max-over-ranks timing reduction, and the `--impl reference` convention that only
rank 0 works.  The GPU path itself shards with no data-path collective."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import bench
    dist.init_process_group("gloo", rank=rank, world_size=world)
    inp = bench.synthetic_inputs(1, 32, 48, 4100, seed=bench.shard_seed(rank))
    ms = bench.max_over_ranks(10.0 + 5.0 * rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, float(inp["pcs"].sum()))
    with open(os.path.join(out_dir, "r%d.json" % rank), "w") as f:
        json.dump({"ms": ms, "sums": gathered}, f)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_and_max_reduce(tmp_path):
    world, port = 2, 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    res = [json.load(open(tmp_path / ("r%d.json" % r))) for r in range(world)]
    assert all(r["ms"] == 15.0 for r in res)                      # max over ranks, identical everywhere
    assert res[0]["sums"] == res[1]["sums"] and res[0]["sums"][0] != res[0]["sums"][1]   # distinct shards


def test_reference_arm_only_rank0_works():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_reference_arm_json_contract():
    """Rank 0 of the reference arm prints one JSON line with the agreed keys (tiny workload)."""
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "small",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "pairs/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0
