"""Throughput sweep of BASELINE.json configs[2..4]: {5,10,20}-way x {1,5}-shot x {8,16,32} frames on ViT-B/16 and the
ViT-L/14 16-frame case, one GPU, device-resident synthetic episodes, random weights (fsar_set_weight from device
tensors). Prints one JSON line per configuration; results are committed under profiles/."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clip_fsar_b200 import lib as L, synth

CASES = [c for c in [("ViT-B/16", 5, 1, 8), ("ViT-B/16", 5, 5, 8), ("ViT-B/16", 5, 1, 16), ("ViT-B/16", 5, 1, 32), ("ViT-B/16", 10, 1, 8),
         ("ViT-B/16", 20, 1, 8), ("ViT-B/16", 10, 5, 16), ("ViT-B/16", 20, 5, 32), ("ViT-L/14", 5, 1, 16)]
         if not os.environ.get("SWEEP_ONLY") or os.environ["SWEEP_ONLY"] == c[0]]

def main():
    dev = torch.device("cuda", 0)
    for geom, way, shot, T in CASES:
        g = synth.full_geometry(geom)
        S, Q = way * shot, way
        frames = (S + Q) * T
        merge = shot > 1
        eng = L.Engine(**dict(g, max_frames=96, max_videos=S + Q, max_tokens=T, max_classes=64, otam_lambda=0.5, device=0))
        gen = torch.Generator(device=dev).manual_seed(0)
        for name, shape in synth.state_dict_shapes(g).items():
            fan_in = shape[-1] if len(shape) == 2 else (shape[1] * shape[2] * shape[3] if len(shape) == 4 else 1)
            w = torch.randn(shape, device=dev, generator=gen) * (fan_in ** -0.5 if len(shape) > 1 else 0.02)
            if name.endswith("norm.weight") or (".ln_" in name and name.endswith("weight")) or name == "scale":
                w = torch.ones(shape, device=dev)
            eng.set_weight(name, w)
        eng.set_weight("text_features_train", torch.randn(64, g["embed_dim"], device=dev, generator=gen))
        eng.set_weight("text_features_test", torch.randn(24, g["embed_dim"], device=dev, generator=gen))
        sup = torch.randn(S * T, 3, 224, 224, device=dev, generator=gen)
        tgt = torch.randn(Q * T, 3, 224, 224, device=dev, generator=gen)
        sl = torch.arange(way, device=dev, dtype=torch.float32).repeat_interleave(shot)[torch.randperm(S, device=dev, generator=gen)].contiguous()
        rl = (sl + 3).contiguous()
        args = (sup, tgt, sl, rl, T, way, merge, False)
        for _ in range(2):
            logits, _ = eng.episode_forward(*args, n_train_classes=64)
        torch.cuda.synchronize()
        n = max(3, min(200, int(16000 / frames)))
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n):
            logits, _ = eng.episode_forward(*args, n_train_classes=64)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        flops = frames * synth.vit_flops_per_frame(g)     # reference-equivalent FLOPs (every row of every block)
        eng.profile_begin()
        eng.episode_forward(*args, n_train_classes=64)
        prof = {k: round(v["ms"], 3) for k, v in eng.profile_end().items() if v["ms"] > 0}
        print(json.dumps(dict(kernel_ms=prof, backbone=geom, way=way, shot=shot, frames_per_video=T, frames_per_episode=frames, merge_before=merge,
                              ms_per_episode=ms, episodes_per_s=1000.0 / ms, vit_tflops=flops / ms / 1e9, finite=bool(torch.isfinite(logits).all()),
                              us_per_frame=ms * 1e3 / frames)), flush=True)
        eng.close(); del sup, tgt; torch.cuda.empty_cache()

if __name__ == "__main__":
    main()
