"""GPU: the reference's UNMODIFIED runner (runs/run.py -> runs/test_net_few_shot.py:test_few_shot / test_epoch, 35-224)
drives the registered sm_100a head on a B200 — `NUM_GPUS: 1`, and `NUM_GPUS: 2` / `8` through the reference's own
`torch.multiprocessing.spawn` launcher + DistributedDataParallel wrap (utils/launcher.py:29-34, models/base/builder.py:69-79)
when the box has that many GPUs. The tree is the byte-identical copy staged by tools/stage_reference.sh under baseline/_ref
(git-ignored, travels with the gpurun snapshot); only YAML files are added next to the config they inherit from.

Checked: the run finishes, the logged `val_epoch` top1_err equals what clip_fsar_b200.runner.evaluate computes in this
process on the same seeded episodes with the same seeded model, and the through-runner episodes/s is written to
gpurun_out/ for profiles/."""
import json
import os
import re
import socket
import subprocess
import sys
import time

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "baseline", "_ref")
CFG_DIR = os.path.join("configs", "projects", "CLIPFSAR", "kinetics100")
N_EPISODES = 24


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_reference_runner(n_gpus, out_dir, n_episodes=N_EPISODES, pool=0):
    name = "tmp_runner_%dgpu_%d.yaml" % (n_gpus, n_episodes)
    with open(os.path.join(REF, CFG_DIR, name), "w") as f:
        f.write("_BASE: ./CLIPFSAR_synth_5way1shot_vitb16_sm100.yaml\n"
                "TRAIN:\n  NUM_TEST_TASKS: %d\n  BATCH_SIZE: %d\n"
                "TEST:\n  BATCH_SIZE: %d\n"
                "DATA_LOADER:\n  NUM_WORKERS: 4\n"
                "LOG_PERIOD: 1\nNUM_GPUS: %d\nOUTPUT_DIR: %s\n" % (n_episodes, n_gpus, n_gpus, n_gpus, out_dir))
    env = dict(os.environ, CLIP_FSAR_ROOT=REF, PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""),
               FSAR_SYNTH_POOL=str(pool))
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, "-m", "clip_fsar_b200.run", "--cfg", os.path.join(CFG_DIR, name),
                        "--init_method", "tcp://127.0.0.1:%d" % _free_port()],
                       capture_output=True, text=True, env=env, timeout=900, cwd=ROOT)
    wall = time.perf_counter() - t0
    assert r.returncode == 0, (r.stdout[-3000:], r.stderr[-3000:])
    text = r.stdout
    for fn in os.listdir(out_dir):
        if fn.endswith(".log"):
            text += open(os.path.join(out_dir, fn)).read()
    epochs = [json.loads(m) for m in re.findall(r'(\{[^{}]*"_type": "val_epoch"[^{}]*\})', text)]
    iters = [json.loads(m) for m in re.findall(r'(\{[^{}]*"_type": "val_iter"[^{}]*\})', text)]
    assert epochs, text[-3000:]
    return epochs[-1], iters, wall


def _expected_top1_err(n_episodes):
    """Same seeded model (torch.manual_seed(cfg.RANDOM_SEED) before build_model, as test_few_shot does) and the same seeded
    episodes (Synth_few_shot: seed 1000 + index) through clip_fsar_b200.runner.evaluate in this process."""
    from clip_fsar_b200 import runner
    from clip_fsar_b200.register import register
    register(REF)
    cwd, argv = os.getcwd(), sys.argv
    os.chdir(REF)
    sys.argv = ["runs/run.py", "--cfg", os.path.join(CFG_DIR, "CLIPFSAR_synth_5way1shot_vitb16_sm100.yaml")]
    try:
        from utils.config import Config
        cfg = Config(load=True)
        cfg.NUM_GPUS = 1
        torch.manual_seed(cfg.RANDOM_SEED)
        from models.base.builder import build_model
        model, _ = build_model(cfg)
    finally:
        os.chdir(cwd)
        sys.argv = argv
    model.eval()
    with torch.no_grad():
        res = runner.evaluate(lambda task: model(task)["logits"], n_episodes, way=5, shot=1, queries_per_class=1, n_frames=8,
                              image_size=224, n_test_classes=len(cfg.TEST.CLASS_NAME), seed=1000, device="cuda:0",
                              structured=True, engine=None)
    return res


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "runs")), reason="reference tree not staged (tools/stage_reference.sh)")
@pytest.mark.parametrize("n_gpus", [1, 2, 8])
def test_unmodified_reference_runner_drives_the_sm100_head(n_gpus, tmp_path):
    if torch.cuda.device_count() < n_gpus:
        pytest.skip("needs %d GPUs" % n_gpus)
    epoch, iters, wall = _run_reference_runner(n_gpus, str(tmp_path))
    want = _expected_top1_err(N_EPISODES)
    assert want["n_total"] == N_EPISODES * 5
    # ValMeter averages the per-iteration error rates (equal-sized episodes): the global error rate
    assert abs(float(epoch["top1_err"]) - want["top1_err"]) < 1e-3, (epoch, want)
    # through-runner rate (SURVEY.md 8d timing protocol): a second, longer run over a pool of cached episodes so that the
    # loader does not generate random frames per index; the runner's own per-iteration timer (utils/meters.py:787), first
    # iterations dropped. Still inside every iteration: collation of 48 MB, .cuda(), the forward, 3 x .item() and
    # (NUM_GPUS > 1) 3 all-reduces (test_net_few_shot.py:59-62, 168-178).
    n_long = 160 * n_gpus
    _, iters_long, wall_long = _run_reference_runner(n_gpus, str(tmp_path), n_episodes=n_long, pool=8)
    dts = sorted(float(i["time_diff"]) for i in iters_long[10:])
    rec = {"n_gpus": n_gpus, "episodes": N_EPISODES, "top1_err_runner": float(epoch["top1_err"]), "top1_err_expected": want["top1_err"],
           "wall_s_whole_run": wall, "throughput_run_episodes": n_long, "throughput_run_wall_s": wall_long,
           "runner_episodes_per_s_median_iter": (n_gpus / dts[len(dts) // 2]) if dts else None,
           "runner_episodes_per_s_mean_iter": (n_gpus * len(dts) / sum(dts)) if dts else None,
           "note": "runs/test_net_few_shot.py:test_epoch unmodified, head = CNN_OTAM_CLIPFSAR_SM100 (one episode per call)"}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r2_reference_runner_%dgpu.json" % n_gpus), "w") as f:
        json.dump(rec, f)
