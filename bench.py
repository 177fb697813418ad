"""bench.py — episodes/s of the CLIP-FSAR few-shot inference path on N B200s (one process per GPU).

Workload (BASELINE.json `metric`, configs[1]): 5-way 1-shot, 1 query per class, 8 frames of 224x224, ViT-B/16,
random-init weights, synthetic frames. A step = one episode (80 frames) through the hot path:
CLIP ViT frame encoder -> temporal prototype modulator -> cosine/OTAM head -> logits.

  value : whole-job episodes/s with the inputs already resident in HBM (a pool of distinct episodes larger than
          L2 is cycled, so no step re-reads its inputs from cache); device-timed, max over ranks. Episodes go
          through fsar_episodes_forward --batch at a time (default 6 = 480 frames = five 96-frame ViT passes, which
          makes every GEMM a whole number of tile waves); --batch 1 gives the one-episode-per-call figure.
  e2e   : the same metric through the C-ABI host entry points (fsar_episodes_submit_host / collect_host): HOST
          pinned buffers in, logits on the host out, H2D + D2H inside the timed region, copies of call i+1
          overlapped with the compute of call i (two slots).
  roofline     : the tcgen05 GEMM kernel (95 % of the FLOPs): algorithmic FLOPs / CUDA-event time of its launches.
  cpu_baseline : the CPU oracle (a port of the reference forward) on this box's host cores, bounded sample.

`--impl reference` times that CPU implementation alone (all host threads) and prints the same line shape.
Launch for N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
                   --master-port P bench.py --gpus N --steps K --warmup W
"""
import argparse
import json
import os
import subprocess
import sys
import time

# torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU-baseline legs (rank 0 only) need all host cores, and the
# OpenMP pool is sized when torch first touches it, so lift the cap before importing torch.
if os.environ.get("OMP_NUM_THREADS", "") in ("", "1"):
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, library chatter)
# is redirected to stderr; the line itself goes to a private duplicate of the original stdout.
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

from clip_fsar_b200 import synth  # noqa: E402

WAY, SHOT, QPC, T, GEOM = 5, 1, 1, 8, "ViT-B/16"
N_TRAIN, N_TEST = 64, 24
METRIC = "episodes/sec (5-way 1-shot, 8x224^2, ViT-B/16)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU, sampled every 100 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            pass

    def start(self):
        time.sleep(0.25)  # let the first sample land before the timed region starts

    def summary(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        rows = [[x.strip() for x in ln.split(",")] for ln in out.strip().splitlines()]
        rows = [r for r in rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi gave no samples"]}
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]),
                "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows), "reasons": reasons}


def cpu_episode_time(sd, g, tt, te, frames_per_sample, budget_s, steps, warmup):
    """Time the CPU oracle (port of the reference forward) on a bounded sample: the ViT on `frames_per_sample` of the
    episode's 80 frames (99.8 % of the CPU time and linear in frames) plus the full head."""
    from oracle import fsar_oracle as O
    task = synth.synth_episode(WAY, SHOT, QPC, T, g["image_size"], N_TEST, 1000, structured=False)
    frames = torch.from_numpy(np.concatenate([task["support_set"], task["target_set"]])[:frames_per_sample])
    E = g["embed_dim"]
    S, Q = WAY * SHOT, WAY * QPC
    feats = torch.randn((S + Q) * T, E)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            f = O.vit_forward(sd, g, frames)
            t_vit = time.perf_counter() - t0
            feats[:f.shape[0]] = f
            t1 = time.perf_counter()
            O.head_forward(sd, g, tt, te, feats[:S * T].reshape(S, T, E), feats[S * T:].reshape(Q, T, E),
                           task["support_labels"], task["real_support_labels"])
            t_head = time.perf_counter() - t1
            if i >= warmup:
                times.append(t_vit * ((S + Q) * T / frames_per_sample) + t_head)
            if sum(times) > budget_s and len(times) >= 1 and i >= warmup:
                break
    return float(np.median(times)), len(times)


def reference_arm(args, rank, world):
    """--impl reference: the reference's CPU forward (oracle port; the reference is Python and cannot travel to the
    GPU box) on all host threads, same config/metric. Rank 0 only."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = synth.full_geometry(GEOM)
    sd = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(g, 0, spread=False).items()}
    tt = synth.synth_text_features(N_TRAIN, g["embed_dim"], 7)
    te = synth.synth_text_features(N_TEST, g["embed_dim"], 8)
    # bounded sample: choose the frame count so that (steps + warmup) samples end within ~150 s
    t_probe, _ = cpu_episode_time(sd, g, tt, te, 8, 1e9, 1, 0)          # extrapolated s / episode from 8 frames
    per_frame = t_probe / 80.0
    budget = 150.0
    fps = 80
    for cand in (80, 40, 16, 8):
        fps = cand
        if per_frame * cand * (args.steps + args.warmup) <= budget:
            break
    t_ep, n = cpu_episode_time(sd, g, tt, te, fps, 1e9, args.steps, args.warmup)
    eps = 1.0 / t_ep
    sample = ("%d timed steps; each step = fp32 CPU forward of %d of the episode's 80 frames through the 12-layer ViT "
              "(extrapolated linearly to 80) + the full modulator/OTAM head" % (n, fps))
    line = {"impl": "reference", "metric": METRIC, "value": eps, "unit": "episodes/s", "n_gpus": args.gpus,
            "steps": n, "warmup": args.warmup, "ms_per_step": t_ep * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "5-way 1-shot, 1 query/class, 8x224^2 frames, ViT-B/16 random-init, 80 frames/episode"},
            "cpu_baseline": {"value": eps, "unit": "episodes/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": eps, "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=6,
                    help="episodes per fsar_episodes_* call; their 80-frame sets are regrouped into 96-frame ViT passes "
                         "(whole waves of 256x256 tiles on 148 SMs). 1 = one episode per call")
    ap.add_argument("--pass-frames", type=int, default=96,
                    help="frames per ViT pass when --batch > 1 (96 x 197 rows = 74 row blocks of 256 = one per CTA pair)")
    ap.add_argument("--pool", type=int, default=6, help="distinct resident episodes cycled (6 x 48 MB > 126 MB L2)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, rank, world)
    if args.warmup < 3:
        args.warmup = 3
    B = max(1, args.batch)
    args.pool = max(args.pool, B)
    args.steps = -(-args.steps // B) * B          # whole calls
    n_calls, n_warm_calls = args.steps // B, -(-args.warmup // B)

    import torch.distributed as dist
    from clip_fsar_b200 import lib as L
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    g = synth.full_geometry(GEOM)
    n_vid = WAY * (SHOT + QPC)
    frames_per_pass = args.pass_frames if B > 1 else n_vid * T
    eng = L.Engine(**dict(g, max_frames=frames_per_pass, max_videos=n_vid, max_tokens=T, max_classes=max(N_TRAIN, N_TEST),
                          max_batch=B, otam_lambda=0.5, device=local))
    sd_np = synth.synth_state_dict(g, 0, spread=False)
    eng.load_state_dict({k: torch.from_numpy(v) for k, v in sd_np.items()})
    tt = synth.synth_text_features(N_TRAIN, g["embed_dim"], 7)
    te = synth.synth_text_features(N_TEST, g["embed_dim"], 8)
    eng.set_weight("text_features_train", torch.from_numpy(tt))
    eng.set_weight("text_features_test", torch.from_numpy(te))
    assert eng.missing_weights() == []

    # independent episodes per rank (weak scaling: every rank runs `steps` episodes of its own)
    keys = ("support_set", "target_set", "support_labels", "real_support_labels")
    host_pool, dev_pool = [], []
    for i in range(args.pool):
        ep = synth.synth_episode(WAY, SHOT, QPC, T, g["image_size"], N_TEST, 1000 + rank * 1_000_000 + i, structured=False)
        host_pool.append([torch.from_numpy(ep[k]).pin_memory() for k in keys])
        dev_pool.append([t.to(dev) for t in host_pool[-1]])
    h2d = sum(t.numel() * 4 for t in host_pool[0])            # per episode (= per step)
    d2h = (WAY * QPC * WAY + n_vid * N_TRAIN) * 4

    def step(i):
        """One call = B episodes (B steps of the metric)."""
        eps = [dev_pool[(i * B + j) % args.pool] for j in range(B)]
        return eng.episodes_forward(eps, T, WAY, n_train_classes=N_TRAIN)

    def host_eps(i):
        return [host_pool[(i * B + j) % args.pool] for j in range(B)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    # ---------------------------------------------------------------- device-resident throughput
    for i in range(n_warm_calls):
        step(i)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    n0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n_calls):
        logits, _ = step(i)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = eng.launch_count() - n0
    clocks = sampler.summary() if sampler else None
    value = args.steps * world / (ms_total / 1e3)

    # ---------------------------------------------------------------- end to end through the host entry points
    out = torch.empty(B, WAY * QPC, WAY)
    cl = torch.empty(B, n_vid, N_TRAIN)
    for i in range(2):
        eng.episodes_submit_host(0, host_eps(i), T, WAY)
        eng.episodes_collect_host(0, out, cl)
    barrier()
    t0 = time.perf_counter()
    eng.episodes_submit_host(0, host_eps(0), T, WAY)
    for i in range(1, n_calls):
        eng.episodes_submit_host(i & 1, host_eps(i), T, WAY)
        eng.episodes_collect_host((i - 1) & 1, out, cl)
    eng.episodes_collect_host((n_calls - 1) & 1, out, cl)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
    barrier()
    e2e_value = args.steps * world / (e2e_ms / 1e3)

    # ---------------------------------------------------------------- per-kernel device time (CUDA events per launch)
    NP = 2 * B if B > 1 else 4     # episodes in the profiled pass
    eng.profile_begin()
    for i in range(NP // B):
        step(i)
    prof = eng.profile_end()
    pk = peaks()
    gemm_classes = [k for k in prof if k.startswith("gemm_")]
    gemm_ms = sum(prof[k]["ms"] for k in gemm_classes)
    gemm_flops = sum(prof[k]["flops"] for k in gemm_classes)
    gemm_launches = sum(prof[k]["launches"] for k in gemm_classes)
    total_prof_ms = sum(v["ms"] for v in prof.values())
    achieved = gemm_flops / gemm_ms / 1e9 if gemm_ms else 0.0
    traffic = None   # DRAM bytes per launch of the GEMM kernel from the committed ncu capture (bench cannot run ncu)
    try:
        nt = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["kernels"]
        gk = [v for k, v in nt.items() if "gemm_tn_tcgen05" in k]
        if gk:
            traffic = sum(v["dram_bytes_per_launch"] * v["launches"] for v in gk) / sum(v["launches"] for v in gk)
    except (OSError, KeyError, ValueError):
        pass
    roofline = {"bound": "tensor", "kernel": "gemm_tn_tcgen05_pair_kernel, full-size launches (patch/QKV/out/fc1/fc2 epilogues over all "
                                             "token rows; the n_frames-row launches of the CLS-only last block are class last_block_cls)",
                "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": achieved / pk["tf_sustained"],
                "peak_source": pk["src"] + " cuBLAS bf16, sustained (kernel timed inside a long step)",
                "traffic": traffic, "traffic_source": "profiles/ncu_traffic.json (ncu --set full, mean over the GEMM launches)",
                "algorithmic_bytes_per_launch": None, "avg_launch_ms": gemm_ms / max(gemm_launches, 1), "launches_per_episode": gemm_launches / NP,
                "share_of_step": gemm_ms / total_prof_ms if total_prof_ms else None,
                "flops_per_episode": gemm_flops / NP}
    kernels = {k: {"ms_per_episode": v["ms"] / NP, "launches_per_episode": v["launches"] / NP,
                   "tflops": (v["flops"] / v["ms"] / 1e9 if v["ms"] and v["flops"] else None),
                   "gbs": (v["bytes"] / v["ms"] / 1e6 if v["ms"] and v["bytes"] else None)} for k, v in prof.items()}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sd = {k: torch.from_numpy(v) for k, v in sd_np.items()}
        t_ep, n = cpu_episode_time(sd, g, tt, te, 16, 20.0, 3, 1)
        cpu = {"value": 1.0 / t_ep, "unit": "episodes/s", "cores": cores, "kind": "port",
               "sample": "%d timed samples of 16 of the episode's 80 frames through the fp32 CPU ViT (extrapolated "
                         "linearly) + the full head, torch CPU ops on %d threads" % (n, cores)}

    # FLOPs: `flops_ref` is the reference-equivalent count (every token row of every block); the library
    # EXECUTES fewer in the last block, where only the CLS row is read downstream. Rates are quoted on
    # executed FLOPs (summed over the launches actually made), never on the skipped ones.
    flops_ref = 80 * synth.vit_flops_per_frame(g)
    flops_ep = sum(prof[k]["flops"] for k in prof if k.startswith("gemm_") or k in ("attention", "final_proj", "last_block_cls")) / NP
    line = {"metric": METRIC, "value": value, "unit": "episodes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16 operands / fp32 accumulate (reference: f32)", "data": "synthetic",
            "config": {"workload": "5-way 1-shot, 1 query/class, 8x224^2 frames, ViT-B/16 random-init, 80 frames/episode",
                       "l2_policy": "inputs larger than L2: %d distinct resident episodes (%.0f MB) cycled" %
                                    (args.pool, args.pool * h2d / 1e6),
                       "episodes_per_call": B, "frames_per_vit_pass": frames_per_pass,
                       "episodes_per_rank": args.steps, "parallelism": "episodes sharded, dp%d, no data-path collective" % world},
            "e2e": {"value": e2e_value, "unit": "episodes/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / args.steps, "api": "fsar_episodes_submit_host/collect_host (2 slots, pinned host buffers, %d episodes per call)" % B},
            "gpu_launches": launches * world, "clocks": clocks,
            "vit_tflops": flops_ep * args.steps * world / (ms_total / 1e3) / 1e12,
            "vit_frac_of_sustained_peak": flops_ep * args.steps / (ms_total / 1e3) / 1e12 / pk["tf_sustained"],
            "vit_flops_per_episode": {"executed": flops_ep, "reference_equivalent": flops_ref,
                                      "note": "last block: Q / out_proj / ln_2 / MLP on the CLS row only (the only row "
                                              "ln_post reads); rates use executed FLOPs"},
            "roofline": roofline, "cpu_baseline": cpu, "kernels": kernels}
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
