"""Frame pre-processing: oracle vs fixtures produced by the reference's own transform objects (CPU),
and the fused CUDA kernel vs the oracle through the C ABI (GPU)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith("preproc_"))


def load(name):
    from clip_fsar_b200 import synth
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    m = json.loads(str(z["meta"]))
    frames = synth.synth_raw_frames(m["T"], m["H"], m["W"], m["seed"])
    assert int(frames.astype(np.int64).sum()) == int(z["frames_checksum"][0])        # the frames the reference saw
    return m, frames, z


@pytest.mark.parametrize("name", CASES)
def test_oracle_preprocess_matches_reference_transforms(name):
    from oracle import fsar_oracle as O
    m, frames, z = load(name)
    out = O.preprocess_u8(frames, m["crop"], (m["scale"], m["scale"]), m["mean"], m["std"]).numpy()
    st = m["stride"]
    assert out.shape == (m["T"], 3, m["crop"], m["crop"])
    assert np.abs(out[:, :, ::st, ::st] - z["out_sub"]).max() < 1e-5
    assert abs(np.float64(out).sum() - z["out_checksum"][0]) < 1e-3 * z["out_checksum"][1] * 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_preprocess_matches_oracle_and_fixture(lib, name):
    from clip_fsar_b200 import synth
    from oracle import fsar_oracle as O
    m, frames, z = load(name)
    g = dict(synth.full_geometry("tiny"), image_size=m["crop"]) if m["crop"] == 32 else synth.full_geometry("ViT-B/16")
    g = dict(g, layers=1)                                                              # weights are not needed here
    e = lib.Engine(**dict(g, max_frames=4, max_videos=2, max_tokens=2, max_classes=4, otam_lambda=0.5, device=0))
    out = e.preprocess_u8(torch.from_numpy(frames).cuda(), (m["scale"], m["scale"]), m["mean"], m["std"]).cpu().numpy()
    ref = O.preprocess_u8(frames, m["crop"], (m["scale"], m["scale"]), m["mean"], m["std"]).numpy()
    # fp32 on both sides; the only difference is the association order of the four bilinear taps
    assert np.abs(out - ref).max() < 2e-5
    st = m["stride"]
    assert np.abs(out[:, :, ::st, ::st] - z["out_sub"]).max() < 2e-5
    e.close()


@pytest.mark.gpu
def test_uint8_host_entry_point_equals_float_path(lib):
    """fsar_episodes_submit_host_u8: raw bytes over PCIe + device pre-processing == pre-processing first (oracle) and
    feeding the fp32 crops to the normal entry point."""
    from conftest import load_golden, regenerate
    from clip_fsar_b200 import synth
    from oracle import fsar_oracle as O
    meta, _ = load_golden("tiny_5w1s")
    g, sd, tt, te, task = regenerate(meta)
    e = lib.Engine(**dict(g, max_frames=80, max_videos=10, max_tokens=8, max_classes=64, max_batch=2, otam_lambda=0.5, device=0))
    e.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    e.set_weight("text_features_train", torch.from_numpy(tt))
    e.set_weight("text_features_test", torch.from_numpy(te))
    eps_u8, eps_f32 = [], []
    for i in range(2):
        sup = synth.synth_raw_frames(40, 48, 64, 100 + i)
        tgt = synth.synth_raw_frames(40, 48, 64, 200 + i)
        sl, rl = torch.from_numpy(task["support_labels"]), torch.from_numpy(task["real_support_labels"])
        eps_u8.append((torch.from_numpy(sup).pin_memory(), torch.from_numpy(tgt).pin_memory(), sl, rl))
        eps_f32.append((O.preprocess_u8(sup, 32, (40, 40)).cuda(), O.preprocess_u8(tgt, 32, (40, 40)).cuda(), sl.cuda(), rl.cuda()))
    want, _ = e.episodes_forward(eps_f32, 8, 5, n_train_classes=64)
    e.episodes_submit_host_u8(0, eps_u8, 8, 5, resize=(40, 40))
    got = torch.empty(2, 5, 5)
    e.episodes_collect_host(0, got)
    # the crops agree to ~1e-6; a few of those flip a 16-bit operand rounding in the ViT, hence not bit equal
    assert float((got - want.cpu()).abs().max() / want.abs().max()) < 1e-3
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("geom,H,W,scale", [("tiny", 48, 64, 40), ("ViT-B/16", 240, 320, 256), ("ViT-B/16", 128, 171, 256),
                                            ("l14-2layer", 360, 640, 256)])
def test_fused_u8_patch_gather_equals_two_pass_route(lib, geom, H, W, scale):
    """SURVEY.md 8f-1 "feeding K1 directly": fsar_vit_forward_u8 evaluates resize / crop / normalise inside the patch gather
    (uint8 THWC -> 16-bit im2col rows, no fp32 NCHW crop in HBM). It must give the frame features of the two-pass route
    fsar_preprocess_u8 -> fsar_vit_forward, whose first half is pinned to the reference's transforms above: same fp32
    arithmetic per pixel, same rounding to the operand type, so the patch rows -- and with them every later number -- are
    bit-identical. Patch sizes 16 and 14 (ViT-L/14), down- and up-sampling sources."""
    from clip_fsar_b200 import synth
    g = dict(synth.full_geometry(geom), layers=1)
    sd = synth.synth_state_dict(g, 5, spread=False)
    e = lib.Engine(**dict(g, max_frames=4, max_videos=2, max_tokens=4, max_classes=4, otam_lambda=0.5, device=0))
    e.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    frames = torch.from_numpy(synth.synth_raw_frames(6, H, W, seed=91)).cuda()          # 6 frames: two passes of <= 4
    two_pass = e.vit_forward(e.preprocess_u8(frames, (scale, scale)))
    fused = e.vit_forward_u8(frames, (scale, scale))
    assert torch.isfinite(fused).all() and float(fused.abs().max()) > 0
    assert torch.equal(fused, two_pass)
    e.close()
