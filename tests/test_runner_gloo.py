"""CPU, world_size 2 over gloo: episode sharding and the single end-of-run counter all-reduce."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _fake_forward(task):
    # deterministic pseudo-logits from the data: enough to exercise sharding + reduction without a GPU
    q = task["target_set"].reshape(task["target_labels"].numel(), -1).mean(1, keepdim=True)
    s = task["support_set"].reshape(task["support_labels"].numel(), -1).mean(1)[None, :]
    order = torch.argsort(task["support_labels"])
    return -(q - s[:, order]).abs()


def _worker(rank, world, port, n_episodes, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from clip_fsar_b200 import runner
    res = runner.evaluate(_fake_forward, n_episodes, n_frames=2, image_size=8, rank=rank, world=world)
    res["mine"] = list(runner.shard(n_episodes, rank, world))
    out[rank] = res
    dist.barrier()
    dist.destroy_process_group()


def test_shard_partition_is_exact():
    from clip_fsar_b200 import runner
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in runner.shard(37, r, world))
        assert seen == list(range(37))


@pytest.mark.timeout(120)
def test_world2_gloo_equals_single_process():
    from clip_fsar_b200 import runner
    n = 7
    single = runner.evaluate(_fake_forward, n, n_frames=2, image_size=8)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29611, n, out), nprocs=2, join=True)
    assert sorted(out[0]["mine"] + out[1]["mine"]) == list(range(n))
    for r in (0, 1):
        for k in ("n_correct", "n_total", "top1_acc"):
            assert out[r][k] == single[k]
        assert abs(out[r]["loss"] - single["loss"]) < 1e-5
    assert single["n_total"] == n * 5


def test_counters_are_integer_and_sync_free():
    from clip_fsar_b200 import runner
    c = runner.new_counters("cpu")
    logits = torch.tensor([[2.0, 0.0], [0.0, 1.0], [3.0, 0.0]])
    runner.update_counters(c, logits, torch.tensor([0.0, 1.0, 1.0]))
    s = runner.summarise(c)
    assert c.dtype == torch.int64 and s["n_correct"] == 2 and s["n_total"] == 3
    ref = torch.nn.functional.cross_entropy(logits, torch.tensor([0, 1, 1])).item()
    assert abs(s["loss"] - ref) < 1e-5
