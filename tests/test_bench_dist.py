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


class _TinyFlow(torch.nn.Module):
    """Stand-in with the contract trainer.train_step needs: forward(inputs) sets `.loss`."""

    def __init__(self):
        super().__init__()
        self.lin = torch.nn.Linear(4, 2)

    def forward(self, inputs):
        self.loss = (self.lin(inputs["x"]) - inputs["flow_2d"]).pow(2).mean()
        return {}


def _train_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from camliflow_b200 import trainer
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = trainer.wrap_ddp(_TinyFlow())
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    g = torch.Generator().manual_seed(100 + rank)                       # every rank owns different "pairs"
    inputs = {"x": torch.randn(8, 4, generator=g), "flow_2d": torch.randn(8, 2, generator=g)}
    loss = trainer.train_step(model, opt, inputs)
    torch.save({"w": trainer.unwrap(model).lin.weight.detach(), "g": trainer.unwrap(model).lin.weight.grad, "loss": loss,
                "x": inputs["x"], "y": inputs["flow_2d"]}, os.path.join(out_dir, "t%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_training_step_allreduces_gradients(tmp_path):
    """trainer.train_step under DistributedDataParallel: both ranks end with identical weights, and the applied
    gradient is the mean of the per-rank gradients (the all-reduce of BASELINE config 5's data-parallel step)."""
    world, port = 2, 31500 + os.getpid() % 2000
    mp.spawn(_train_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / ("t%d.pt" % i)) for i in range(world)]
    assert torch.equal(r[0]["w"], r[1]["w"]) and torch.equal(r[0]["g"], r[1]["g"])
    torch.manual_seed(0)
    ref = _TinyFlow()
    grads = []
    for i in range(world):
        ref.zero_grad()
        ref({"x": r[i]["x"], "flow_2d": r[i]["y"]})
        ref.loss.backward()
        grads.append(ref.lin.weight.grad.clone())
    assert torch.allclose(r[0]["g"], (grads[0] + grads[1]) / 2, atol=1e-6)


def _flat_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from camliflow_b200 import trainer
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = _TinyFlow().train()
    g = torch.Generator().manual_seed(100 + rank)
    inputs = {"x": torch.randn(8, 4, generator=g), "flow_2d": torch.randn(8, 2, generator=g)}
    step = trainer.CapturedTrainStep(model, inputs, lr=0.1, weight_decay=0.0, max_grad_norm=None, use_graph=False)
    loss = step(inputs)
    torch.save({"w": model.lin.weight.detach().clone(), "g": model.lin.weight.grad.clone(), "loss": loss.clone(),
                "flat": step.flat.clone(), "x": inputs["x"], "y": inputs["flow_2d"]}, os.path.join(out_dir, "f%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_flat_allreduce_step(tmp_path):
    """trainer.CapturedTrainStep (eager mode on CPU): every parameter's gradient is a view of ONE flat buffer, the
    buffer is mean-reduced over the ranks with a single all-reduce, and both ranks apply the same update."""
    world, port = 2, 33500 + os.getpid() % 2000
    mp.spawn(_flat_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / ("f%d.pt" % i)) for i in range(world)]
    assert torch.equal(r[0]["w"], r[1]["w"]) and torch.equal(r[0]["flat"], r[1]["flat"])
    assert r[0]["flat"].numel() == 4 * 2 + 2                     # weight + bias of the tiny model, one buffer
    torch.manual_seed(0)
    ref = _TinyFlow()
    grads = []
    for i in range(world):
        ref.zero_grad()
        ref({"x": r[i]["x"], "flow_2d": r[i]["y"]})
        ref.loss.backward()
        grads.append(ref.lin.weight.grad.clone())
    assert torch.allclose(r[0]["g"], (grads[0] + grads[1]) / 2, atol=1e-6)
    assert r[0]["loss"] != r[1]["loss"]                          # (each rank reports the loss of its own pairs)
