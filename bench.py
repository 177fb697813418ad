"""bench.py — episodes/s of the CLIP-FSAR few-shot inference path on N B200s (one process per GPU).

A step = one episode through the hot path: CLIP ViT frame encoder -> temporal prototype modulator -> cosine/OTAM head
-> logits. `--workload` picks the configuration (BASELINE.json `configs`):

  headline (default, configs[1], the one `metric` is quoted on): 5-way 1-shot, 1 query/class, 8 x 224^2, ViT-B/16
  5w5s     (configs[2]) : 5-way 5-shot, MERGE_BEFORE prototypes, 8 frames, ViT-B/16 (240 frames / episode)
  l14_t16  (configs[3]) : 5-way 1-shot, 16 frames, ViT-L/14 (160 frames / episode)
  sweep    (configs[4]) : {5,10,20}-way x {1,5}-shot x {8,16,32} frames, ViT-B/16; one row per point with ep/s, the
                          fraction of the tensor roofline and a parity bit against the fp16-operand-emulating oracle

What one JSON line carries (rank 0 prints it, stdout holds nothing else):
  value  : whole-job episodes/s, inputs resident in HBM (a pool of distinct episodes larger than L2 is cycled), device
           timed (CUDA events), max over ranks. The timed loop runs max(--steps, what fills --min-seconds) episodes,
           rounded up to whole calls of `episodes_per_call`; `steps` is the number actually timed.
  module_path : the same metric one episode per call (fsar_episode_forward), i.e. what the registered nn.Module and the
           reference runner can reach.
  e2e    : through the C-ABI host entry points: pinned HOST fp32 frames in, logits on the host out, H2D + D2H inside
           the timed region, two slots. `e2e_u8`: the same with RAW uint8 frames over PCIe (fsar_episodes_submit_host_u8).
  roofline : the tcgen05 GEMM kernel, CUDA events around every launch in a pass taken right AFTER the sustained region
           with its own nvidia-smi clock record; both denominators (burst / sustained) are printed and `frac` uses the
           one that matches the observed clock state.
  counters : top-1 hits / queries / cross-entropy accumulated on the device every call (fsar_metrics_update) and summed
           over ranks by ONE NCCL all_reduce(int64[3]) at the end -- the collective that replaces
           runs/test_net_few_shot.py:168-171.
  cpu_baseline : the reference's own forward on this box's host cores (unmodified tree staged under baseline/_ref, else
           the oracle port), bounded sample.

`--impl reference` times that CPU implementation alone (all host threads) and prints the same line shape.
Launch for N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
                   --master-port P bench.py --gpus N --steps K --warmup W
"""
import argparse
import json
import math
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

N_TRAIN, N_TEST = 64, 24
# the unmodified reference tree staged by tools/stage_reference.sh (git-ignored, travels to the GPU box with the snapshot)
REF_ROOT = os.environ.get("FSAR_REF_ROOT") or os.path.join(ROOT, "baseline", "_ref")
_SD_CACHE = {}


def state_dict_np(geom):
    if geom not in _SD_CACHE:
        _SD_CACHE[geom] = synth.synth_state_dict(synth.full_geometry(geom), 0, spread=False)
    return _SD_CACHE[geom]

WORKLOADS = {
    "headline": dict(geom="ViT-B/16", way=5, shot=1, qpc=1, T=8, merge=False, batch=6, pass_frames=96,
                     metric="episodes/sec (5-way 1-shot, 8x224^2, ViT-B/16)",
                     desc="5-way 1-shot, 1 query/class, 8x224^2 frames, ViT-B/16 random-init, 80 frames/episode"),
    "5w5s": dict(geom="ViT-B/16", way=5, shot=5, qpc=1, T=8, merge=True, batch=2, pass_frames=96,
                 metric="episodes/sec (5-way 5-shot, 8x224^2, ViT-B/16 + temporal prototype modulator)",
                 desc="5-way 5-shot (MERGE_BEFORE), 1 query/class, 8x224^2 frames, ViT-B/16 random-init, 240 frames/episode"),
    # 73 frames x 257 tokens = 18761 rows = 74 row blocks of 256: one per CTA pair
    "l14_t16": dict(geom="ViT-L/14", way=5, shot=1, qpc=1, T=16, merge=False, batch=5, pass_frames=73,
                    metric="episodes/sec (5-way 1-shot, 16x224^2, ViT-L/14)",
                    desc="5-way 1-shot, 1 query/class, 16x224^2 frames, ViT-L/14 random-init, 160 frames/episode"),
}
SWEEP = [(w, s, t) for w in (5, 10, 20) for s in (1, 5) for t in (8, 16, 32)]


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    src="MEASURED_PEAKS.json (cuBLAS bf16 8192^3)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback of B200_PROFILING.md")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU, sampled every 100 ms while a timed region runs."""
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
        rows = rows[2:] if len(rows) > 4 else rows            # the first samples predate the load
        sm = sorted(float(r[0]) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][1]),
                "power_w_max": max(float(r[2]) for r in rows), "samples": len(rows), "reasons": reasons}


def bind_to_gpu_numa_node(index):
    """Run this rank (and allocate its pinned staging buffers: first touch) on the CPUs of the NUMA node its GPU hangs
    off, so the H2D copies of the e2e path do not cross the socket interconnect. Returns what was done."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if not bus:
            return {"bound": False, "why": "no pci.bus_id"}
        bus = bus[-12:] if len(bus) > 12 else bus           # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return {"bound": False, "why": "single NUMA node"}
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"bound": False, "why": "node %d has no allowed CPUs" % node}
        os.sched_setaffinity(0, cpus)
        return {"bound": True, "node": node, "cpus": len(cpus)}
    except (OSError, ValueError, subprocess.SubprocessError) as e:
        return {"bound": False, "why": "%s" % e}


def unbind_cpus():
    """Give every thread of this process all host CPUs again (the CPU-baseline leg uses them all)."""
    every = set(range(os.cpu_count() or 1))
    try:
        for tid in os.listdir("/proc/self/task"):
            try:
                os.sched_setaffinity(int(tid), every)
            except OSError:
                pass
    except OSError:
        pass


# ------------------------------------------------------------------------------------------------ CPU arms
def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "models", "base"))


class RealReference:
    """The UNMODIFIED reference (baseline/_ref, staged by tools/stage_reference.sh) run through its own public API:
    BaseVideoModel(cfg)(task_dict) under model.eval() / torch.no_grad() on the host cores (ipdb /
    ftfy stubs, CLIP checkpoint download replaced by a random-init CLIP of the same geometry, `.cuda()` = identity)."""

    def __init__(self, wl):
        from oracle import gen_golden as G            # test infrastructure: only the CPU arms may use it
        self.G = G
        self.wl = wl
        g = synth.full_geometry(wl["geom"])
        cwd = os.getcwd()
        fs, BaseVideoModel = G.import_reference(REF_ROOT)
        os.chdir(cwd)
        with G.cpu_forward():
            self.model = G.build_reference(fs, BaseVideoModel, g, N_TRAIN, N_TEST, wl["T"], dict(merge_before=wl["merge"]))
        sd = state_dict_np(wl["geom"])
        self.model.head.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
        self.model.head.text_features_train = torch.from_numpy(synth.synth_text_features(N_TRAIN, g["embed_dim"], 7))
        self.model.head.text_features_test = torch.from_numpy(synth.synth_text_features(N_TEST, g["embed_dim"], 8))
        self.g = g

    def time_episode(self, way):
        """Seconds for one forward of a `way`-way sub-episode of the workload (same shot / queries / frames)."""
        wl = self.wl
        ep = synth.synth_episode(way, wl["shot"], wl["qpc"], wl["T"], self.g["image_size"], N_TEST, 1000, structured=False)
        task = {k: torch.from_numpy(v) for k, v in ep.items()}
        with self.G.cpu_forward(), torch.no_grad():
            t0 = time.perf_counter()
            out = self.model(task)
            dt = time.perf_counter() - t0
        assert out["logits"].shape == (way * wl["qpc"], way)
        return dt


def port_episode_time(wl, frames_per_sample):
    """The oracle port (oracle/fsar_oracle.py) on a bounded sample: the ViT on `frames_per_sample` frames (99.8 % of the CPU
    time and linear in frames), extrapolated to the episode, plus the full head."""
    from oracle import fsar_oracle as O
    g = synth.full_geometry(wl["geom"])
    sd = {k: torch.from_numpy(v) for k, v in state_dict_np(wl["geom"]).items()}
    tt = synth.synth_text_features(N_TRAIN, g["embed_dim"], 7)
    te = synth.synth_text_features(N_TEST, g["embed_dim"], 8)
    task = synth.synth_episode(wl["way"], wl["shot"], wl["qpc"], wl["T"], g["image_size"], N_TEST, 1000, structured=False)
    S, Q, T, E = wl["way"] * wl["shot"], wl["way"] * wl["qpc"], wl["T"], g["embed_dim"]
    frames = torch.from_numpy(np.concatenate([task["support_set"], task["target_set"]])[:frames_per_sample])
    feats = torch.randn((S + Q) * T, E)
    with torch.no_grad():
        t0 = time.perf_counter()
        f = O.vit_forward(sd, g, frames)
        t_vit = time.perf_counter() - t0
        feats[:f.shape[0]] = f
        t1 = time.perf_counter()
        O.head_forward(sd, g, tt, te, feats[:S * T].reshape(S, T, E), feats[S * T:].reshape(Q, T, E),
                       task["support_labels"], task["real_support_labels"], merge_before=wl["merge"])
        t_head = time.perf_counter() - t1
    return t_vit * ((S + Q) * T / frames_per_sample) + t_head


def cpu_arm(wl, steps, warmup, budget_s):
    """Time the CPU implementation of the path for `steps` steps after `warmup`, each step a bounded sample sized so the
    whole run fits `budget_s`. Returns (seconds per FULL episode (median), n timed, kind, sample description)."""
    cores = os.cpu_count() or 1
    unbind_cpus()
    torch.set_num_threads(cores)
    frames_full = wl["way"] * (wl["shot"] + wl["qpc"]) * wl["T"]
    n = max(1, steps) + max(0, warmup)
    if reference_available():
        try:
            ref = RealReference(wl)
            t1 = ref.time_episode(1)                                  # probe: a 1-way sub-episode
            per_frame = t1 / ((wl["shot"] + wl["qpc"]) * wl["T"])
            way = wl["way"]
            while way > 1 and per_frame * way * (wl["shot"] + wl["qpc"]) * wl["T"] * n > budget_s:
                way -= 1
            frames = way * (wl["shot"] + wl["qpc"]) * wl["T"]
            times = []
            for i in range(n):
                dt = ref.time_episode(way)
                if i >= warmup:
                    times.append(dt * frames_full / frames)
            sample = ("%d timed steps; each step = the unmodified reference's BaseVideoModel(cfg)(task_dict) (eval, no_grad, fp32, "
                      "%d host threads) on a %d-way sub-episode = %d of the episode's %d frames%s" %
                      (len(times), cores, way, frames, frames_full,
                       "" if way == wl["way"] else ", scaled linearly in frames (the ViT is 99.8 % of the CPU time)"))
            return float(np.median(times)), len(times), "reference", sample
        except Exception as e:                                        # noqa: BLE001  (fall back to the port, say why)
            sys.stderr.write("reference arm: the staged reference failed (%s: %s); timing the oracle port\n" % (type(e).__name__, e))
    t8 = port_episode_time(wl, 8)
    fps = frames_full
    for cand in (frames_full, frames_full // 2, 40, 16, 8):
        fps = max(8, min(cand, frames_full))
        if t8 * fps / frames_full * n <= budget_s:
            break
    times = []
    for i in range(n):
        dt = port_episode_time(wl, fps)
        if i >= warmup:
            times.append(dt)
    sample = ("%d timed steps; each step = the oracle port's fp32 CPU forward of %d of the episode's %d frames through the ViT "
              "(extrapolated linearly) + the full modulator/OTAM head, %d host threads" % (len(times), fps, frames_full, cores))
    return float(np.median(times)), len(times), "port", sample


def reference_arm(args, rank):
    """--impl reference: rank 0 alone runs and prints; the other ranks exit 0 without work."""
    if rank != 0:
        return
    wl = WORKLOADS["headline" if args.workload == "sweep" else args.workload]
    t_ep, n, kind, sample = cpu_arm(wl, args.steps, args.warmup, 150.0)
    eps = 1.0 / t_ep
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": wl["metric"], "value": eps, "unit": "episodes/s", "n_gpus": args.gpus,
            "steps": n, "warmup": args.warmup, "ms_per_step": t_ep * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": wl["desc"]},
            "cpu_baseline": {"value": eps, "unit": "episodes/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": eps, "unit": "episodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=_JSON_OUT, flush=True)


# ------------------------------------------------------------------------------------------------ the B200 arm
class Dist:
    def __init__(self, rank, world, dev):
        self.rank, self.world, self.dev = rank, world, dev

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max(self, x):
        if self.world > 1:
            import torch.distributed as dist
            t = torch.tensor([x], device=self.dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x


class Workload:
    """Engine + resident / pinned episode pools of one workload on this rank."""

    def __init__(self, wl, L, dev, local, rank, batch=None, pass_frames=None, pool=None, want_host=True, want_u8=True):
        self.wl, self.L, self.dev = wl, L, dev
        g = self.g = synth.full_geometry(wl["geom"])
        self.way, self.T = wl["way"], wl["T"]
        self.S, self.Q = wl["way"] * wl["shot"], wl["way"] * wl["qpc"]
        self.n_vid = self.S + self.Q
        self.frames = self.n_vid * self.T
        self.B = B = max(1, batch or wl["batch"])
        self.pass_frames = pass_frames or wl["pass_frames"]
        frame_bytes = 3 * g["image_size"] ** 2 * 4
        # distinct resident episodes: more than L2 (126 MB) and at least one call's worth
        self.pool = max(pool or 0, B, -(-160_000_000 // (self.frames * frame_bytes)))
        self.eng = L.Engine(**dict(g, max_frames=self.pass_frames if (B > 1 or self.frames > self.pass_frames) else self.frames,
                                   max_videos=self.n_vid, max_tokens=self.T, max_classes=max(N_TRAIN, N_TEST), max_batch=B,
                                   otam_lambda=0.5, device=local))
        self.sd_np = state_dict_np(wl["geom"])
        self.eng.load_state_dict({k: torch.from_numpy(v) for k, v in self.sd_np.items()})
        self.tt = synth.synth_text_features(N_TRAIN, g["embed_dim"], 7)
        self.te = synth.synth_text_features(N_TEST, g["embed_dim"], 8)
        self.eng.set_weight("text_features_train", torch.from_numpy(self.tt))
        self.eng.set_weight("text_features_test", torch.from_numpy(self.te))
        assert self.eng.missing_weights() == []
        keys = ("support_set", "target_set", "support_labels", "real_support_labels")
        self.tasks, self.host_pool, self.dev_pool, self.tgt_labels, self.u8_pool = [], [], [], [], []
        for i in range(self.pool):
            ep = synth.synth_episode(wl["way"], wl["shot"], wl["qpc"], self.T, g["image_size"], N_TEST,
                                     1000 + rank * 1_000_000 + i, structured=False)
            if i == 0:
                self.tasks.append(ep)
            host = [torch.from_numpy(ep[k]) for k in keys]
            self.dev_pool.append([t.to(dev) for t in host])
            self.tgt_labels.append(torch.from_numpy(ep["target_labels"]).to(dev))
            if want_host:
                self.host_pool.append([t.pin_memory() for t in host])
            if want_u8:
                rng = np.random.default_rng(77 + i)
                S_img = g["image_size"]
                self.u8_pool.append([torch.from_numpy(rng.integers(0, 256, size=(n * self.T, S_img, S_img, 3), dtype=np.uint8)).pin_memory()
                                     for n in (self.S, self.Q)] + [t.pin_memory() for t in host[2:]])
        self.h2d = sum(t.numel() * 4 for t in (self.host_pool[0] if want_host else self.dev_pool[0]))
        self.h2d_u8 = (sum(t.numel() for t in self.u8_pool[0][:2]) + 8 * self.S) if want_u8 else None
        self.d2h = (self.Q * self.way + self.n_vid * N_TRAIN) * 4
        self.counters = torch.zeros(3, dtype=torch.int64, device=dev)

    def close(self):
        self.eng.close()
        self.dev_pool = self.host_pool = self.u8_pool = None
        torch.cuda.empty_cache()

    # one call = B episodes (B steps of the metric); the metric counters are updated on the device, no host sync
    def call(self, i, count=True):
        B, pool = self.B, self.pool
        eps = [self.dev_pool[(i * B + j) % pool] for j in range(B)]
        logits, cl = self.eng.episodes_forward(eps, self.T, self.way, merge_before=self.wl["merge"], n_train_classes=N_TRAIN)
        if count:
            for j in range(B):
                self.eng.metrics_update(logits[j], self.tgt_labels[(i * B + j) % pool], self.counters)
        return logits, cl

    def timed_device(self, n_calls, dist, sampler=None):
        dist.barrier()
        if sampler:
            sampler.start()
        n0 = self.eng.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_calls):
            self.call(i)
        e1.record()
        dist.barrier()
        ms = dist.max(e0.elapsed_time(e1))
        return ms, self.eng.launch_count() - n0, (sampler.summary() if sampler else None)

    def timed_host(self, n_calls, dist, u8=False):
        """End to end: pinned HOST buffers in, HOST logits out, two slots so the copies of call i+1 overlap call i."""
        B, pool, eng = self.B, self.pool, self.eng
        out = torch.empty(B, self.Q, self.way)
        cl = torch.empty(B, self.n_vid, N_TRAIN)
        src = self.u8_pool if u8 else self.host_pool
        S_img = self.g["image_size"]

        def submit(slot, i):
            eps = [src[(i * B + j) % pool] for j in range(B)]
            if u8:
                eng.episodes_submit_host_u8(slot, eps, self.T, self.way, resize=(S_img, S_img), merge_before=self.wl["merge"])
            else:
                eng.episodes_submit_host(slot, eps, self.T, self.way, merge_before=self.wl["merge"])

        for i in range(2):
            submit(0, i)
            eng.episodes_collect_host(0, out, cl)
        dist.barrier()
        t0 = time.perf_counter()
        submit(0, 0)
        for i in range(1, n_calls):
            submit(i & 1, i)
            eng.episodes_collect_host((i - 1) & 1, out, cl)
        eng.episodes_collect_host((n_calls - 1) & 1, out, cl)
        torch.cuda.synchronize()
        ms = dist.max((time.perf_counter() - t0) * 1e3)
        dist.barrier()
        return ms

    def profile(self, n_calls, sampler=None):
        if sampler:
            sampler.start()
        self.eng.profile_begin()
        for i in range(n_calls):
            self.call(i, count=False)
        prof = self.eng.profile_end()
        return prof, (sampler.summary() if sampler else None)

    def parity(self, full_frames_limit):
        """Parity bit of this workload on ONE episode against the fp16-operand-emulating oracle (CPU, rank 0)."""
        from oracle import fsar_oracle as O
        ep, wl, g = self.tasks[0], self.wl, self.g
        sd = {k: torch.from_numpy(v) for k, v in self.sd_np.items()}
        d = self.dev_pool[0]
        logits, _ = self.eng.episode_forward(d[0], d[1], d[2], d[3], self.T, self.way, merge_before=wl["merge"],
                                             n_train_classes=N_TRAIN)
        got = logits.cpu()
        E = g["embed_dim"]
        if self.frames <= full_frames_limit:
            ref = O.episode_forward(sd, g, self.tt, self.te, ep, self.T, wl["merge"], False, operand_dtype=self.eng.operand_dtype)
            want, mode = ref["logits"], "full episode"
            vit_err = None
        else:
            # large episodes: the head oracle on the library's frame features + the ViT oracle on 8 sampled frames
            sf = self.eng.peek("support_feats", (self.S, self.T, E))
            tf = self.eng.peek("target_feats", (self.Q, self.T, E))
            ref = O.head_forward(sd, g, self.tt, self.te, sf, tf, ep["support_labels"], ep["real_support_labels"],
                                 merge_before=wl["merge"])
            want, mode = ref["logits"], "head oracle on device features + ViT oracle on 8 sampled frames"
            idx = np.linspace(0, self.S * self.T - 1, 8).astype(int)
            f16 = O.vit_forward(sd, g, torch.from_numpy(ep["support_set"][idx]), operand_dtype=self.eng.operand_dtype)
            vit_err = float((sf.reshape(-1, E)[idx] - f16).norm() / f16.norm())
        rel = float((got - want).abs().max() / want.abs().max())
        ok = rel < 1e-3 and (vit_err is None or vit_err < 2e-3)
        return {"ok": bool(ok), "logits_max_rel": rel, "vit_feature_rel_l2": vit_err, "mode": mode, "bound": 1e-3}


def gemm_roofline(prof, n_episodes, pk, clocks, step_clocks=None):
    classes = [k for k in prof if k.startswith("gemm_")]
    ms = sum(prof[k]["ms"] for k in classes)
    flops = sum(prof[k]["flops"] for k in classes)
    byts = sum(prof[k]["bytes"] for k in classes)
    launches = sum(prof[k]["launches"] for k in classes)
    total_ms = sum(v["ms"] for v in prof.values())
    achieved = flops / ms / 1e9 if ms else 0.0
    # which cuBLAS figure is the fair denominator: the burst one if the clock stayed at its maximum with no power cap
    # during the pass, else the sustained one (MEASURED_PEAKS.json holds both); both fractions are printed
    capped = True
    if clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz"):
        capped = ("sw_power_cap" in clocks.get("reasons", [])) or clocks["sm_mhz"] < 0.97 * clocks["sm_max_mhz"]
    peak = pk["tf_sustained"] if capped else pk["tf_burst"]
    traffic = None
    try:
        nt = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["kernels"]
        gk = [v for k, v in nt.items() if "gemm_tn_tcgen05" in k]
        if gk:
            traffic = sum(v["dram_bytes_per_launch"] * v["launches"] for v in gk) / sum(v["launches"] for v in gk)
    except (OSError, KeyError, ValueError):
        pass
    return {"bound": "tensor",
            "kernel": "gemm_tn_tcgen05_pair_kernel, full-size launches (patch/QKV/out/fc1/fc2 epilogues over all token rows; the "
                      "n_frames-row launches of the CLS-only last block are class last_block_cls)",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
            "peak_kind": "sustained (power-capped / clock below max during the pass)" if capped else "burst (clock at max, no power cap)",
            "frac_of_sustained": achieved / pk["tf_sustained"], "frac_of_burst": achieved / pk["tf_burst"],
            "peak_sustained": pk["tf_sustained"], "peak_burst": pk["tf_burst"], "peak_source": pk["src"],
            "clocks": clocks, "launches": launches, "avg_launch_ms": ms / max(launches, 1),
            "flops_per_launch": flops / max(launches, 1), "algorithmic_bytes_per_launch": byts / max(launches, 1),
            "traffic": traffic, "traffic_source": "profiles/ncu_traffic.json (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum, mean over the GEMM launches)",
            "launches_per_episode": launches / n_episodes, "share_of_step": ms / total_ms if total_ms else None,
            "flops_per_episode": flops / n_episodes,
            # the event-bracketed pass leaves gaps between kernels, draws less power and therefore clocks higher than the
            # back-to-back sustained region: the same rate scaled to the SM clock observed THERE (tensor-bound kernel)
            "achieved_at_step_clock": (achieved * step_clocks["sm_mhz"] / clocks["sm_mhz"]
                                       if step_clocks and clocks and step_clocks.get("sm_mhz") and clocks.get("sm_mhz") else None),
            "frac_at_step_clock": (achieved * step_clocks["sm_mhz"] / clocks["sm_mhz"] / pk["tf_sustained"]
                                   if step_clocks and clocks and step_clocks.get("sm_mhz") and clocks.get("sm_mhz") else None),
            "how": "CUDA events around every launch (fsar_profile_begin/end) over %d episodes right after the sustained region; "
                   "achieved = flops_per_launch / avg_launch_ms; frac_at_step_clock = achieved x (SM clock of the sustained "
                   "region / SM clock of this pass) / sustained peak" % n_episodes}


def executed_flops(prof, n_episodes):
    return sum(prof[k]["flops"] for k in prof
               if k.startswith("gemm_") or k in ("attention", "final_proj", "last_block_cls")) / n_episodes


def run_workload(name, args, L, dev, local, dist):
    wl = WORKLOADS[name]
    rank, world = dist.rank, dist.world
    W = Workload(wl, L, dev, local, rank, batch=args.batch, pass_frames=args.pass_frames, pool=args.pool)
    B = W.B
    pk = peaks()
    warm_calls = max(1, -(-max(args.warmup, 3) // B))
    for i in range(warm_calls):
        W.call(i, count=False)
    # size the timed region: max(--steps, what fills --min-seconds), whole calls; every rank agrees on the count
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    W.call(0, count=False)
    e1.record()
    torch.cuda.synchronize()
    call_ms = dist.max(e0.elapsed_time(e1))
    want = args.steps if args.exact_steps else max(args.steps, int(math.ceil(args.min_seconds * 1e3 / call_ms * B)))
    n_calls = int(dist.max(float(-(-want // B))))
    steps = n_calls * B
    W.counters.zero_()
    ms_total, launches, clocks = W.timed_device(n_calls, dist, ClockSampler(local) if rank == 0 else None)
    value = steps * world / (ms_total / 1e3)
    # the one collective of the path: accuracy counters, summed once over NCCL (runs/test_net_few_shot.py:168-171)
    counters = W.counters.clone()
    if world > 1:
        import torch.distributed as tdist
        tdist.all_reduce(counters, op=tdist.ReduceOp.SUM)
    c = counters.tolist()
    counters_out = {"n_correct": c[0], "n_total": c[1], "loss_sum": c[2] / 1e6, "expected_total": steps * world * W.Q,
                    "ok": c[1] == steps * world * W.Q,
                    "how": "fsar_metrics_update after every call (device int64 counters, no host sync); one %s at the end"
                           % ("all_reduce(int64[3], SUM) over NCCL" if world > 1 else "read (single rank: no collective)")}
    assert counters_out["ok"], counters_out

    # per-kernel pass right after the sustained region, with its own clock record
    prof_calls = max(2, int(math.ceil(0.6e3 / call_ms)))
    prof, prof_clocks = W.profile(prof_calls, ClockSampler(local) if rank == 0 else None)
    NP = prof_calls * B
    roofline = gemm_roofline(prof, NP, pk, prof_clocks, clocks)
    kernels = {k: {"ms_per_episode": v["ms"] / NP, "launches_per_episode": v["launches"] / NP,
                   "tflops": (v["flops"] / v["ms"] / 1e9 if v["ms"] and v["flops"] else None),
                   "gbs": (v["bytes"] / v["ms"] / 1e6 if v["ms"] and v["bytes"] else None)} for k, v in prof.items()}
    flops_ep = executed_flops(prof, NP)

    e2e = e2e_u8 = module = None
    if not args.no_extras:
        e2e_calls = max(3, n_calls // 2)
        ms = W.timed_host(e2e_calls, dist)
        e2e = {"value": e2e_calls * B * world / (ms / 1e3), "unit": "episodes/s", "h2d_bytes_per_step": W.h2d,
               "d2h_bytes_per_step": W.d2h, "ms_per_step": ms / (e2e_calls * B), "steps": e2e_calls * B,
               "api": "fsar_episodes_submit_host/collect_host (2 slots, pinned host fp32 frames, %d episodes per call)" % B}
        ms = W.timed_host(e2e_calls, dist, u8=True)
        e2e_u8 = {"value": e2e_calls * B * world / (ms / 1e3), "unit": "episodes/s", "h2d_bytes_per_step": W.h2d_u8,
                  "d2h_bytes_per_step": W.d2h, "ms_per_step": ms / (e2e_calls * B), "steps": e2e_calls * B,
                  "api": "fsar_episodes_submit_host_u8/collect_host (raw uint8 THWC frames over PCIe, resize/crop/normalise on the "
                         "device; different pixel values than the fp32 pool, same shapes)"}
    parity = W.parity(args.parity_frames) if (rank == 0 and args.parity) else None
    frames_ep = W.frames
    W.close()

    if not args.no_extras:
        # the path the registered nn.Module / the reference runner uses: one episode per call
        M = Workload(wl, L, dev, local, rank, batch=1, pass_frames=min(W.pass_frames, frames_ep), want_host=False, want_u8=False)
        for i in range(3):
            M.call(i, count=False)
        m_calls = max(10, int(math.ceil(0.7e3 * B / call_ms)))
        ms, m_launches, _ = M.timed_device(m_calls, dist)
        module = {"value": m_calls * world / (ms / 1e3), "unit": "episodes/s", "steps": m_calls, "ms_per_step": ms / m_calls,
                  "launches_per_episode": m_launches / m_calls,
                  "api": "fsar_episode_forward, one episode per call (CNN_OTAM_CLIPFSAR_SM100.forward)"}
        module["vs_batched"] = module["value"] / value
        M.close()

    if rank != 0:
        return None
    flops_ref = frames_ep * synth.vit_flops_per_frame(W.g)
    line = {"metric": wl["metric"], "value": value, "unit": "episodes/s", "n_gpus": world, "steps": steps,
            "warmup": warm_calls * B, "ms_per_step": ms_total / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp16 operands / fp32 accumulate (reference: f32)", "data": "synthetic",
            "config": {"workload": wl["desc"], "workload_key": name,
                       "l2_policy": "inputs larger than L2: %d distinct resident episodes (%.0f MB) cycled" % (W.pool, W.pool * W.h2d / 1e6),
                       "episodes_per_call": B, "frames_per_vit_pass": W.pass_frames, "requested_steps": args.steps,
                       "min_seconds": None if args.exact_steps else args.min_seconds, "timed_seconds": ms_total / 1e3,
                       "episodes_per_rank": steps, "parallelism": "episodes sharded, dp%d, no data-path collective" % world},
            "e2e": e2e, "e2e_u8": e2e_u8, "module_path": module, "counters": counters_out,
            "gpu_launches": launches * world, "clocks": clocks,
            "vit_tflops": flops_ep * steps * world / (ms_total / 1e3) / 1e12,
            "vit_frac_of_sustained_peak": flops_ep * steps / (ms_total / 1e3) / 1e12 / pk["tf_sustained"],
            "vit_frac_of_burst_peak": flops_ep * steps / (ms_total / 1e3) / 1e12 / pk["tf_burst"],
            "vit_flops_per_episode": {"executed": flops_ep, "reference_equivalent": flops_ref,
                                      "note": "last block: Q / out_proj / ln_2 / MLP on the CLS row only (the only row ln_post "
                                              "reads); rates use executed FLOPs"},
            "roofline": roofline, "parity": parity, "kernels": kernels}
    return line


def run_sweep(args, L, dev, local, dist):
    """configs[4]: one row per (way, shot, T) point; every rank runs the point on its own episodes (weak scaling)."""
    rank, world = dist.rank, dist.world
    pk = peaks()
    rows = []
    points = SWEEP if not args.sweep_points else [tuple(int(x) for x in p.split("x")) for p in args.sweep_points.split(",")]
    for way, shot, T in points:
        frames = way * (shot + 1) * T
        # episodes per call: the count (<= 6) that fills 96-frame passes best
        B = min(range(1, 7), key=lambda b: (-(-b * frames // 96) * 96 / (b * frames), b))
        wl = dict(geom="ViT-B/16", way=way, shot=shot, qpc=1, T=T, merge=shot > 1, batch=B, pass_frames=96, metric="", desc="")
        W = Workload(wl, L, dev, local, rank, want_host=False, want_u8=False)
        for i in range(2):
            W.call(i, count=False)
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        W.call(0, count=False)
        e1.record()
        torch.cuda.synchronize()
        call_ms = dist.max(e0.elapsed_time(e1))
        n_calls = int(dist.max(float(max(2, math.ceil(args.sweep_seconds * 1e3 / call_ms)))))
        ms, launches, clocks = W.timed_device(n_calls, dist, ClockSampler(local) if rank == 0 else None)
        prof, _ = W.profile(1)
        flops_ep = executed_flops(prof, B)
        eps = n_calls * B * world / (ms / 1e3)
        tfl = flops_ep * n_calls * B / (ms / 1e3) / 1e12            # per GPU
        row = {"way": way, "shot": shot, "frames_per_video": T, "frames_per_episode": frames, "merge_before": shot > 1,
               "episodes_per_call": B, "steps": n_calls * B, "episodes_per_s": eps, "ms_per_episode": ms / (n_calls * B),
               "vit_tflops_per_gpu": tfl, "frac": tfl / pk["tf_sustained"], "frac_of_burst": tfl / pk["tf_burst"],
               "sm_mhz": clocks and clocks.get("sm_mhz"), "reasons": clocks and clocks.get("reasons"),
               "gpu_launches": launches * world,
               "parity": W.parity(args.parity_frames) if (rank == 0 and args.parity) else None}
        W.close()
        rows.append(row)
        if rank == 0:
            sys.stderr.write("sweep %2d-way %d-shot T=%2d: %8.2f ep/s, frac %.3f, parity %s\n" %
                             (way, shot, T, eps, row["frac"], row["parity"] and row["parity"]["ok"]))
    if rank != 0:
        return None
    head = next((r for r in rows if (r["way"], r["shot"], r["frames_per_video"]) == (5, 1, 8)), rows[0])
    return {"metric": "episodes/sec, sweep {5,10,20}-way x {1,5}-shot x {8,16,32} frames (ViT-B/16); value = the 5-way 1-shot 8-frame point",
            "value": head["episodes_per_s"], "unit": "episodes/s", "n_gpus": world, "steps": head["steps"], "warmup": 2,
            "ms_per_step": head["ms_per_episode"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp16 operands / fp32 accumulate (reference: f32)", "data": "synthetic",
            "config": {"workload": "sweep: {5,10,20}-way x {1,5}-shot x {8,16,32} frames, 224^2, ViT-B/16 random-init, 1 query/class; "
                                   "5-shot points use MERGE_BEFORE prototypes", "workload_key": "sweep",
                       "frac": "executed ViT FLOPs / time / cuBLAS bf16 sustained peak (%.1f TFLOP/s), per GPU" % pk["tf_sustained"],
                       "l2_policy": "inputs larger than L2 (>= 160 MB of distinct resident frames cycled per point)",
                       "seconds_per_point": args.sweep_seconds, "parallelism": "episodes sharded, dp%d" % world},
            "gpu_launches": sum(r["gpu_launches"] for r in rows),
            "parity_all_ok": all(r["parity"]["ok"] for r in rows) if args.parity else None, "sweep": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=6)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="headline", choices=sorted(WORKLOADS) + ["sweep"])
    ap.add_argument("--min-seconds", type=float, default=1.5,
                    help="the timed region runs at least this long (a sustained, power-capped figure); `steps` reports what was timed")
    ap.add_argument("--exact-steps", action="store_true", help="time exactly --steps (rounded up to whole calls), however short")
    ap.add_argument("--batch", type=int, default=None, help="episodes per fsar_episodes_* call (default: per workload; 1 = one per call)")
    ap.add_argument("--pass-frames", type=int, default=None, help="frames per ViT pass (default: per workload)")
    ap.add_argument("--pool", type=int, default=None, help="distinct resident episodes cycled (default: > 160 MB of frames)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip e2e / e2e_u8 / module_path (kernel work only)")
    ap.add_argument("--no-parity", dest="parity", action="store_false", help="skip the oracle parity bit")
    ap.add_argument("--parity-frames", type=int, default=320, help="episodes up to this many frames get the full-episode oracle")
    ap.add_argument("--sweep-seconds", type=float, default=0.6)
    ap.add_argument("--sweep-points", default=None, help="e.g. 5x1x8,20x5x32 (way x shot x frames)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, rank)

    import torch.distributed as tdist
    from clip_fsar_b200 import lib as L
    numa = bind_to_gpu_numa_node(local)          # before any pinned allocation (first touch decides the node)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        tdist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    dist = Dist(rank, world, dev)

    if args.workload == "sweep":
        line = run_sweep(args, L, dev, local, dist)
    else:
        line = run_workload(args.workload, args, L, dev, local, dist)
    if rank == 0:
        line["numa"] = numa
        if args.workload != "sweep" and not args.no_cpu_baseline:
            wl = WORKLOADS[args.workload]
            t_ep, n, kind, sample = cpu_arm(wl, 3, 1, 30.0)
            line["cpu_baseline"] = {"value": 1.0 / t_ep, "unit": "episodes/s", "cores": os.cpu_count() or 1, "kind": kind,
                                    "sample": sample}
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
