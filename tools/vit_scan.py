import sys, os, json, time, torch
sys.path.insert(0, os.getcwd())
from clip_fsar_b200 import lib as L, synth
g = synth.full_geometry("ViT-B/16")
sd = synth.synth_state_dict(g, 0, spread=False)
for nf in (80, 96, 88, 112, 160, 192):
    eng = L.Engine(**dict(g, max_frames=nf, max_videos=24, max_tokens=8, max_classes=64, otam_lambda=0.5, device=0))
    eng.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    x = torch.randn(nf, 3, 224, 224, device="cuda")
    for _ in range(5): eng.vit_forward(x)
    torch.cuda.synchronize()
    t0 = time.time(); n = 0
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    while time.time() - t0 < 2.0:
        for _ in range(20): eng.vit_forward(x)
        n += 20; torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(json.dumps(dict(frames=nf, ms=ms, us_per_frame=ms * 1e3 / nf, eps_equiv=1000.0 / (ms * 80 / nf))), flush=True)
    eng.close()
