"""Sustained loops of the four ViT GEMM shapes (96-frame pass): TFLOP/s, SM clock, watts. With FSAR_GEMM_DEBUG=1/2 the
epilogue is cut short (results wrong) to see what bounds the main loop."""
import json, os, subprocess, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the -DFSAR_PROBES build (python clip_fsar_b200/build.py --probes) honours FSAR_GEMM_DEBUG / FSAR_ATT_DEBUG
os.environ.setdefault("FSAR_LIB_PATH", os.path.join(ROOT, "clip_fsar_b200", "libfsar_sm100_probes.so"))
from clip_fsar_b200 import lib as L, synth

def sample_start():
    return subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                            stdout=subprocess.PIPE, text=True)
def sample_stop(p):
    p.terminate(); out, _ = p.communicate(timeout=5)
    rows = [[float(x) for x in l.split(",")] for l in out.strip().splitlines() if "," in l]
    rows = rows[3:] if len(rows) > 6 else rows
    return (sorted(r[0] for r in rows)[len(rows) // 2], max(r[1] for r in rows)) if rows else (0, 0)

def run(name, fn, flops, secs=1.0):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    p = sample_start(); time.sleep(0.3)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    n = 0; t0 = time.time(); e0.record()
    while time.time() - t0 < secs:
        for _ in range(50): fn()
        n += 50
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    mhz, wmax = sample_stop(p)
    util = flops / ms / 1e9 / (mhz * 1e6 * 148 * 8192 / 1e12) if mhz else None
    print(json.dumps(dict(kernel=name, dbg=os.environ.get("FSAR_GEMM_DEBUG", "0"), us=round(ms * 1e3, 1), tflops=round(flops / ms / 1e9, 1), sm_mhz=mhz, w_max=wmax,
                          tensor_util=round(util, 3) if util else None)), flush=True)
    time.sleep(0.5)

g = synth.full_geometry("tiny")
eng = L.Engine(**dict(g, max_frames=16, max_videos=10, max_tokens=8, max_classes=64, otam_lambda=0.5, device=0))
dt = eng.operand_dtype; DEV = "cuda:0"; M = int(os.environ.get("PROBE_M", 96 * 197))
for (N, K, epi, nm) in () if os.environ.get("PROBE_ONLY") == "att" else ((2304, 768, L.EPI_STORE16, "qkv"), (768, 768, L.EPI_RESID32, "out"), (3072, 768, L.EPI_QGELU16, "fc1"), (768, 3072, L.EPI_RESID32, "fc2")):
    a = (torch.randn(M, K, device=DEV) * 0.1).to(dt); w = (torch.randn(N, K, device=DEV) * 0.1).to(dt); b = torch.randn(N, device=DEV)
    out = torch.zeros(M, N, device=DEV, dtype=dt if epi in (0, 1) else torch.float32)
    run("gemm_" + nm, lambda: eng.op_gemm(a, w, b, epi, out=out), 2.0 * M * N * K)
if os.environ.get("PROBE_ONLY") == "gemm": sys.exit(0)
qkv = torch.randn(M, 2304, device=DEV).to(dt)
run("attention", lambda: eng.op_attention(qkv, M // 197, 197, 12), 4.0 * (M // 197) * 12 * 197 * 197 * 64)
qkv = torch.randn(96 * 257, 3 * 1024, device=DEV).to(dt)
run("attention_l14", lambda: eng.op_attention(qkv, 96, 257, 16), 4.0 * 96 * 16 * 257 * 257 * 64)
