"""Episodic evaluation sharded across ranks (one process per GPU).

Mirrors what runs/test_net_few_shot.py:test_epoch (35-224) measures — top-1 accuracy and mean cross-entropy over
independent N-way K-shot episodes — with the B200 communication plan of SURVEY.md 8e: episode i goes to rank
i mod world (what DistributedSampler does, datasets/base/builder.py:40-43), counters stay ON DEVICE for the whole
run, and ONE all-reduce of int64[3] = [n_correct, n_total, round(loss_sum * 1e6)] replaces the reference's three
4-byte all-reduces + three .item() syncs per episode (test_net_few_shot.py:168-178). No data-path collective.
"""
import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import synth

LOSS_SCALE = 1_000_000


def shard(n_episodes, rank, world):
    """Global episode indices of this rank."""
    return range(rank, n_episodes, world)


def new_counters(device):
    return torch.zeros(3, dtype=torch.int64, device=device)


def update_counters(counters, logits, target_labels):
    """Device-side: no host synchronisation. logits [Q, way] fp32, target_labels [Q] (fp32 holding integers)."""
    tgt = target_labels.long()
    counters[0] += (logits.argmax(dim=1) == tgt).sum()
    counters[1] += tgt.numel()
    # the reference divides by BATCH_SIZE (=1 episode per rank per step) at test_net_few_shot.py:111
    counters[2] += torch.round(F.cross_entropy(logits, tgt, reduction="sum").double() * LOSS_SCALE).long()
    return counters


def update_counters_device(engine, counters, logits, target_labels, per_class=None):
    """Same accumulation as update_counters, as ONE device kernel through the C ABI (fsar_metrics_update)."""
    return engine.metrics_update(logits, target_labels, counters, per_class)


def reduce_counters(counters):
    """The only collective of the path: one SUM all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def summarise(counters):
    c = counters.tolist()
    total = max(c[1], 1)
    return {"n_correct": c[0], "n_total": c[1], "top1_acc": c[0] / total, "top1_err": 100.0 * (1 - c[0] / total),
            "loss": c[2] / LOSS_SCALE / total}


def evaluate(forward_fn, n_episodes, way=5, shot=1, queries_per_class=1, n_frames=8, image_size=224,
             n_test_classes=24, seed=1000, rank=0, world=1, device="cpu", structured=True, engine=None):
    """Run this rank's shard of `n_episodes` seeded synthetic episodes through `forward_fn(task) -> logits` and
    return the globally reduced summary (identical on every rank). With `engine` the counters are updated by the
    library's device kernel, otherwise by the equivalent torch ops (CPU tests)."""
    counters = new_counters(device)
    for i in shard(n_episodes, rank, world):
        ep = synth.synth_episode(way, shot, queries_per_class, n_frames, image_size, n_test_classes, seed + i, structured)
        task = {k: torch.from_numpy(v).to(device, non_blocking=True) for k, v in ep.items()}
        logits = forward_fn(task)
        if engine is not None:
            update_counters_device(engine, counters, logits, task["target_labels"])
        else:
            update_counters(counters, logits, task["target_labels"])
    return summarise(reduce_counters(counters))
