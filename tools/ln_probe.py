"""LayerNorm over the 96-frame residual stream (18912 x 768 fp32 -> fp16): L2-warm (same buffer back to back: 58 + 29 MB fit the
126 MB L2) vs cold (eight buffers in rotation). Tells how much of the in-situ 17-18 us per launch is DRAM."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clip_fsar_b200 import lib as L, synth

g = synth.full_geometry("tiny")
eng = L.Engine(**dict(g, max_frames=16, max_videos=10, max_tokens=8, max_classes=64, otam_lambda=0.5, device=0))
M, D = 96 * 197, 768
gam, bet = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
bufs = [torch.randn(M, D, device="cuda") for _ in range(8)]

def run(name, pick, n=300):
    for i in range(20): eng.op_layernorm(bufs[pick(i)], gam, bet, True)
    torch.cuda.synchronize()
    eng.profile_begin()                      # CUDA events around every launch: device time, not the host's launch rate
    for i in range(n): eng.op_layernorm(bufs[pick(i)], gam, bet, True)
    prof = eng.profile_end()["layernorm"]
    us = prof["ms"] / prof["launches"] * 1e3
    print(json.dumps(dict(case=name, us=round(us, 2), tb_s=round(M * D * 6 / us / 1e6, 2))), flush=True)

run("warm (one buffer)", lambda i: 0)
run("cold (8 buffers)", lambda i: i % 8)
run("warm (one buffer)", lambda i: 0)
run("cold (8 buffers)", lambda i: i % 8)
