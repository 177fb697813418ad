"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run here (the container that has /root/reference); the GPU box only sees the committed fixtures:
    python oracle/gen_golden.py [--only NAME]

What is real reference code: models.base.few_shot.{CLIP, VisionTransformer, Transformer_v1, cos_sim,
OTAM_cum_dist_v2, CNN_OTAM_CLIPFSAR.forward} and models.base.models.BaseVideoModel, imported as they lie.
What is patched: `ipdb` / `ftfy` stubs; `few_shot.load` returns a random-init CLIP of the requested
geometry instead of downloading a checkpoint; `Tensor.cuda` is the identity on this CPU-only box; the seeded
synthetic state_dict (clip_fsar_b200/synth.py) is loaded with strict=True (which also pins the parameter names);
text_features_{train,test} (plain attributes, few_shot.py:2720/2728) are overwritten with seeded arrays because
the text tower is init-time only and out of scope; for non-ViT-B/16 geometries `context2` is rebuilt with the
reference's own Transformer_v1 class at the right width (the reference hard-codes mid_dim = 512, few_shot.py:2713).
"""
import argparse
import json
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from clip_fsar_b200 import synth  # noqa: E402

CASES = {
    # name: geometry, way, shot, T, flags, weight seed, episode seed
    "tiny_5w1s": dict(geom="tiny", way=5, shot=1, T=8),
    "tiny_5w5s_merge": dict(geom="tiny", way=5, shot=5, T=8, merge_before=True),
    "tiny_5w5s_nomerge": dict(geom="tiny", way=5, shot=5, T=8),
    "tiny_10w1s_T16": dict(geom="tiny", way=10, shot=1, T=16),
    "tiny_3w2s_T32_single": dict(geom="tiny", way=3, shot=2, T=32, single_direct=True),
    "tiny_5w1s_depth2": dict(geom="tiny", way=5, shot=1, T=8, mod_depth=2),
    "tiny_5w1s_default_init": dict(geom="tiny", way=5, shot=1, T=8, spread=False, structured=False),
    "small_5w1s": dict(geom="small", way=5, shot=1, T=8),
    "vitb16_5w1s": dict(geom="ViT-B/16", way=5, shot=1, T=8),
    # largest corner of BASELINE.json's sweep: 20-way 5-shot, 32 frames (100 support + 20 query videos, 3840 frames)
    "tiny_20w5s_T32_merge": dict(geom="tiny", way=20, shot=5, T=32, merge_before=True, slim=True),
    # unequal shots per class: 3 / 1 / 2 support videos (class means over whatever members exist, few_shot.py:2949-2962)
    "tiny_ragged_3w": dict(geom="tiny", way=3, shot=3, T=8, keep_counts=[3, 1, 2]),
    "tiny_ragged_3w_merge": dict(geom="tiny", way=3, shot=3, T=8, keep_counts=[1, 3, 2], merge_before=True),
    # several queries per class (QUERY_PER_CLASS_TEST > 1): 15 query videos against 5 prototypes
    "tiny_5w1s_q3": dict(geom="tiny", way=5, shot=1, T=8, qpc=3),
    # text branches of the eval forward (few_shot.py:2835-2930)
    "tiny_5w5s_evaltext": dict(geom="tiny", way=5, shot=5, T=8, eval_text=True),
    "tiny_5w1s_combine": dict(geom="tiny", way=5, shot=1, T=8, combine=True),
    "tiny_5w5s_combine_coff05_merge": dict(geom="tiny", way=5, shot=5, T=8, combine=True, text_coff=0.5, merge_before=True),
    # the configuration BASELINE.json names: ViT-B/16 "random-init" (default-style init, unstructured N(0,1) frames)
    "vitb16_5w1s_default_init": dict(geom="ViT-B/16", way=5, shot=1, T=8, spread=False, structured=False, eseed=1001),
    # BASELINE.json configs[3]: ViT-L/14 at FULL depth (24 layers, 257 tokens, width 1024), 5-way 1-shot, 16 frames =
    # 160 frames through the reference's own VisionTransformer / Transformer_v1 / OTAM code (the reference head has no
    # ViT-L/14 branch, few_shot.py:2705-2713: mid_dim / context2 are re-built at width 768 with the reference's own
    # classes, see build_reference). emu16: the fp16-operand-emulating oracle's outputs are stored next to the
    # reference's so the GPU test needs no 26-TFLOP CPU forward on the box.
    "vitl14_5w1s_T16_default_init": dict(geom="ViT-L/14", way=5, shot=1, T=16, spread=False, structured=False, eseed=1002,
                                         emu16=True),
}
# Every flag variant again with DEFAULT-INIT weights and unstructured frames (the configuration north_star's 1e-3 bound
# is stated on); the "spread" fixtures above stay as stress cases. Outputs only (slim), a few KB each.
for _name in ("tiny_5w5s_merge", "tiny_5w5s_nomerge", "tiny_10w1s_T16", "tiny_3w2s_T32_single", "tiny_5w1s_depth2",
              "tiny_20w5s_T32_merge", "tiny_ragged_3w", "tiny_ragged_3w_merge", "tiny_5w1s_q3", "tiny_5w5s_evaltext",
              "tiny_5w1s_combine", "tiny_5w5s_combine_coff05_merge"):
    CASES[_name + "_di"] = dict(CASES[_name], spread=False, structured=False, slim=True, eseed=1003)


def import_reference(root=None):
    """Import the unmodified reference from `root` (default /root/reference; bench.py passes baseline/_ref, the staged
    copy that travels to the GPU box)."""
    root = root or REF

    def stub(name, **kw):
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m
        return m

    stub("ipdb", set_trace=lambda *a, **k: None)
    stub("ftfy", fix_text=lambda s: s)
    sys.path.insert(0, root)
    import models.base.few_shot as fs
    from models.base.models import BaseVideoModel
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    return fs, BaseVideoModel


class cpu_forward:
    """Context manager: the reference head hard-codes `.cuda()` in its constructor (few_shot.py:2719, 2726). On a GPU
    box the CPU arm of bench.py still has to build and run it on the HOST cores, so `.cuda()` is the identity while the
    reference is being built / called, and restored afterwards (the library's own code never calls `.cuda()`)."""

    def __enter__(self):
        self.saved = torch.Tensor.cuda
        torch.Tensor.cuda = lambda t, *a, **k: t
        return self

    def __exit__(self, *exc):
        torch.Tensor.cuda = self.saved
        return False


def build_reference(fs, BaseVideoModel, g, n_train, n_test, T, flags):
    heads = g["width"] // 64
    assert heads == g["heads"]
    fs.load = lambda name, cfg=None, device="cpu", jit=False, **kw: (
        fs.CLIP(g["embed_dim"], g["image_size"], g["layers"], g["width"], g["patch_size"], 77, 49408, 64, 1, 1)
        .float().eval(), None)
    NS = types.SimpleNamespace
    train = NS(CLASS_NAME=["c%d" % i for i in range(n_train)], WAY=5, SHOT=1, BATCH_SIZE=1)
    if flags.get("merge_before"):
        train.MERGE_BEFORE = True
    if flags.get("single_direct"):
        train.SINGLE_DIRECT = True
    if flags.get("mod_depth", 1) > 1:
        train.TRANSFORMER_DEPTH = flags["mod_depth"]
    if flags.get("eval_text"):
        train.EVAL_TEXT = True
    if flags.get("combine"):
        train.COMBINE = True
    if flags.get("text_coff"):
        train.TEXT_COFF = flags["text_coff"]
    cfg = NS(TRAIN=train, TEST=NS(CLASS_NAME=["t%d" % i for i in range(n_test)]), DATA=NS(NUM_INPUT_FRAMES=T),
             VIDEO=NS(HEAD=NS(NAME="CNN_OTAM_CLIPFSAR", BACKBONE_NAME="ViT-B/16"), BACKBONE=NS(META_ARCH="Identity")),
             BN=NS(FREEZE=False))
    model = BaseVideoModel(cfg)
    head = model.head
    if g["embed_dim"] != 512:
        head.mid_dim = g["embed_dim"]
        head.context2 = fs.Transformer_v1(dim=g["embed_dim"], heads=8, dim_head_k=g["embed_dim"] // 8, dropout_atte=0.2,
                                          depth=flags.get("mod_depth", 1))
    return model.eval()


def run_case(name, fs, BaseVideoModel, out_dir):
    case = dict(CASES[name])
    g = synth.full_geometry(case["geom"], case.get("mod_depth", 1))
    way, shot, T = case["way"], case["shot"], case["T"]
    n_train, n_test = 64, 24
    wseed, eseed = case.get("wseed", 0), case.get("eseed", 1000)
    sd = synth.synth_state_dict(g, seed=wseed, spread=case.get("spread", True))
    text_train = synth.synth_text_features(n_train, g["embed_dim"], seed=7)
    text_test = synth.synth_text_features(n_test, g["embed_dim"], seed=8)
    task = synth.synth_episode(way, shot, case.get("qpc", 1), T, g["image_size"], n_test, seed=eseed,
                               structured=case.get("structured", True))
    if case.get("keep_counts"):
        task = synth.ragged_support(task, T, case["keep_counts"])

    model = build_reference(fs, BaseVideoModel, g, n_train, n_test, T, case)
    head = model.head
    missing = head.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    head.text_features_train = torch.from_numpy(text_train)
    head.text_features_test = torch.from_numpy(text_test)

    taps = {"backbone": [], "context2": [], "dists": []}
    h1 = head.backbone.register_forward_hook(lambda m, i, o: taps["backbone"].append(o.detach().clone()))
    h2 = head.context2.register_forward_hook(lambda m, i, o: taps["context2"].append(o.detach().clone()))
    orig_otam = fs.OTAM_cum_dist_v2

    def rec_otam(d, lbda=0.5):
        taps["dists"].append(d.detach().clone())
        return orig_otam(d, lbda)

    fs.OTAM_cum_dist_v2 = rec_otam
    try:
        with torch.no_grad():
            out = model({k: torch.from_numpy(v) for k, v in task.items()})
    finally:
        fs.OTAM_cum_dist_v2 = orig_otam
        h1.remove()
        h2.remove()

    E = g["embed_dim"]
    meta = dict(case=name, geom=case["geom"], way=way, shot=shot, T=T, n_train=n_train, n_test=n_test, wseed=wseed,
                eseed=eseed, spread=case.get("spread", True), structured=case.get("structured", True),
                merge_before=bool(case.get("merge_before")), single_direct=bool(case.get("single_direct")),
                mod_depth=case.get("mod_depth", 1), text_seeds=[7, 8],
                text_mode=1 if case.get("eval_text") else (2 if case.get("combine") else 0),
                text_coff=case.get("text_coff", 0.9), keep_counts=case.get("keep_counts"), qpc=case.get("qpc", 1), reference_commit="30cf0a8c",
                torch=torch.__version__, state_dict_keys=sorted(head.state_dict().keys()))
    no_ctx = len(taps["context2"]) == 0       # EVAL_TEXT never runs the modulator / OTAM
    arrays = dict(
        logits=out["logits"].numpy(),
        class_logits=out["class_logits"].numpy() if out["class_logits"] is not None else np.zeros((0,), np.float32),
        support_feats=taps["backbone"][0].reshape(-1, T, E).numpy(), target_feats=taps["backbone"][1].reshape(-1, T, E).numpy(),
        target_mod=np.zeros((0,), np.float32) if no_ctx else taps["context2"][0].numpy(),
        support_mod=np.zeros((0,), np.float32) if no_ctx else taps["context2"][1].numpy(),
        dists=np.zeros((0,), np.float32) if no_ctx else taps["dists"][0].numpy(),
        # checksums of the regenerated inputs so a test can tell "generator drifted" from "oracle is wrong"
        weight_checksum=np.array([float(np.sum([np.float64(v).sum() for v in sd.values()]))]),
        input_checksum=np.array([float(np.float64(task["support_set"]).sum() + np.float64(task["target_set"]).sum())]),
    )
    if case.get("emu16"):
        from oracle import fsar_oracle as O
        emu = O.episode_forward(sd, g, text_train, text_test, task, T, bool(case.get("merge_before")),
                                bool(case.get("single_direct")), operand_dtype=torch.float16)
        arrays["emu16_logits"] = emu["logits"].numpy()
        arrays["emu16_support_feats"] = emu["support_feats"].numpy()
        arrays["emu16_target_feats"] = emu["target_feats"].numpy()
    if case.get("slim"):     # big episodes: keep the outputs, drop the MB-sized intermediates
        for k in ("support_feats", "target_feats", "target_mod", "support_mod", "dists"):
            arrays[k] = np.zeros((0,), np.float32)
    path = os.path.join(out_dir, name + ".npz")
    np.savez_compressed(path, meta=np.array(json.dumps(meta)), **arrays)
    lg = arrays["logits"]
    print("%-26s logits mean %.4f row-spread %.4f argmax %s  -> %s (%d KB)" % (
        name, lg.mean(), (lg.max(1) - lg.min(1)).mean(), lg.argmax(1).tolist(), path, os.path.getsize(path) // 1024))


PREPROC_CASES = {
    # name: source H, W, frames, TEST_SCALE (resize), TEST_CROP_SIZE, stored pixel stride
    "preproc_240x320_to_224": dict(H=240, W=320, T=2, scale=256, crop=224, stride=5),
    "preproc_360x640_to_224": dict(H=360, W=640, T=1, scale=256, crop=224, stride=5),
    "preproc_128x171_to_224": dict(H=128, W=171, T=2, scale=256, crop=224, stride=5),      # up-sampling
    "preproc_48x64_to_32": dict(H=48, W=64, T=4, scale=40, crop=32, stride=1),             # the tiny tower's frame size
}


def run_preproc_case(name, out_dir):
    """Frames through the reference's OWN transform objects, composed exactly as datasets/base/ssv2_few_shot.py:615-642
    composes them at test time."""
    import torchvision.transforms._transforms_video as transforms
    from torchvision.transforms import Compose
    from datasets.utils.transformations import KineticsResizedCropFewshot
    c = PREPROC_CASES[name]
    mean, std = [0.48145466, 0.4578275, 0.40821073], [0.26862954, 0.26130258, 0.27577711]
    resize = KineticsResizedCropFewshot(short_side_range=[c["scale"], c["scale"]], crop_size=c["crop"], num_spatial_crops=1, idx=1)
    tf = Compose([transforms.ToTensorVideo(), resize, transforms.NormalizeVideo(mean=mean, std=std, inplace=True)])
    frames = synth.synth_raw_frames(c["T"], c["H"], c["W"], seed=77)
    out = tf(torch.from_numpy(frames)).permute(1, 0, 2, 3).contiguous().numpy()          # [C,T,H,W] -> [T,C,H,W] (get_seq)
    st = c["stride"]
    meta = dict(case=name, seed=77, **c, mean=mean, std=std, reference_commit="30cf0a8c")
    path = os.path.join(out_dir, name + ".npz")
    np.savez_compressed(path, meta=np.array(json.dumps(meta)), out_sub=out[:, :, ::st, ::st],
                        frames_checksum=np.array([np.int64(frames.astype(np.int64).sum())]),
                        out_checksum=np.array([np.float64(out).sum(), np.float64(np.abs(out)).sum()]))
    print("%-28s out %s -> %s (%d KB)" % (name, out.shape, path, os.path.getsize(path) // 1024))


TEXT_CASES = {
    # name: text geometry (clip_fsar_b200/synth.py TEXT_GEOMETRIES), embed_dim, prompt template, class names
    "text_tiny": dict(geom="tiny", embed_dim=128, prompt="a photo of {}",
                      names=["riding a bike", "playing guitar", "jumping", "x", "pouring water into a glass of water slowly"]),
    # the real ViT-B/16 text tower: width 512, 8 heads, 12 layers; SSv2-style and Kinetics-style class names
    "text_vitb16": dict(geom="ViT-B/16", embed_dim=512, prompt="a photo of {}",
                        names=["Pouring something into something", "Pushing something so that it falls off the table",
                               "air drumming", "blasting sand", "busking", "cutting watermelon", "dancing ballet",
                               "diving cliff", "filling eyebrows", "folding paper", "hula hooping", "ice skating",
                               "paragliding", "playing trumpet", "shearing sheep", "unboxing"]),
    # the ViT-L/14 text tower: width 768, 12 heads, 12 layers, embed 768
    "text_vitl14": dict(geom="ViT-L/14", embed_dim=768, prompt="a photo of {}",
                        names=["brush hair", "cartwheel", "catch", "chew", "climb stairs", "fencing"]),
    "text_tiny_prompt": dict(geom="tiny", embed_dim=128, prompt="a video of a person {}, a type of action",
                             names=["stretching arm", "throwing axe", "side kick"]),
}


def run_text_case(name, fs, out_dir):
    """Token ids from the reference's own BPE tokenizer (few_shot.py:393-429, bpe_simple_vocab_16e6.txt.gz) and text
    features from the reference's own CLIP.encode_text (793-806) with the seeded text-tower weights loaded into it."""
    c = TEXT_CASES[name]
    tg = synth.TEXT_GEOMETRIES[c["geom"]]
    E = c["embed_dim"]
    sd = synth.synth_text_state_dict(tg, E, seed=3)
    clip = fs.CLIP(E, 32, 1, 128, 16, tg["context_length"], tg["vocab_size"], tg["width"], tg["heads"], tg["layers"]).float().eval()
    res = clip.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.startswith("visual.") or k == "logit_scale" for k in res.missing_keys), res.missing_keys
    prompts = [c["prompt"].format(nm) for nm in c["names"]]                       # few_shot.py:2715-2718
    tokens = fs.tokenize(prompts)
    with torch.no_grad():
        feats = clip.encode_text(tokens)
    meta = dict(case=name, geom=c["geom"], embed_dim=E, wseed=3, prompts=prompts, reference_commit="30cf0a8c",
                torch=torch.__version__)
    path = os.path.join(out_dir, name + ".npz")
    np.savez_compressed(path, meta=np.array(json.dumps(meta)), tokens=tokens.numpy().astype(np.int32), features=feats.numpy(),
                        weight_checksum=np.array([float(np.sum([np.float64(v).sum() for v in sd.values()]))]))
    cs = torch.nn.functional.normalize(feats, dim=-1)
    off = (cs @ cs.T - torch.eye(len(prompts))).abs().max().item()
    print("%-26s features %s |f| %.3f max off-diagonal cosine %.3f eot %s -> %s (%d KB)" % (
        name, tuple(feats.shape), feats.norm(dim=-1).mean().item(), off, tokens.argmax(-1).tolist()[:6], path,
        os.path.getsize(path) // 1024))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    cwd = os.getcwd()
    fs, BaseVideoModel = import_reference()
    os.chdir(cwd)
    for name in CASES:
        if a.only and name != a.only:
            continue
        run_case(name, fs, BaseVideoModel, a.out)
    for name in PREPROC_CASES:
        if a.only and name != a.only:
            continue
        run_preproc_case(name, a.out)
    for name in TEXT_CASES:
        if a.only and name != a.only:
            continue
        run_text_case(name, fs, a.out)


if __name__ == "__main__":
    main()
