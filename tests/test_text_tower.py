"""CLIP text tower: fixtures hold token ids from the reference's own BPE tokenizer and features from the
reference's own CLIP.encode_text (oracle/gen_golden.py run_text_case). CPU: the oracle against them. GPU: fsar_text_encode
through the C ABI against the fixtures and against the oracle with 16-bit operand emulation."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from clip_fsar_b200 import synth
from oracle import fsar_oracle as O

CASES = ["text_tiny", "text_tiny_prompt", "text_vitb16", "text_vitl14"]


def regenerate_text(meta):
    tg = synth.TEXT_GEOMETRIES[meta["geom"]]
    return tg, synth.synth_text_state_dict(tg, meta["embed_dim"], meta["wseed"])


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-30))


@pytest.mark.parametrize("name", CASES)
def test_oracle_text_encode_matches_reference(name):
    meta, ref = load_golden(name)
    tg, sd = regenerate_text(meta)
    assert np.isclose(sum(np.float64(v).sum() for v in sd.values()), ref["weight_checksum"][0], rtol=0, atol=1e-6)
    out = O.text_encode(sd, tg, ref["tokens"]).numpy()
    assert out.shape == ref["features"].shape
    assert rel(out, ref["features"]) < 2e-5            # fp32 vs fp32: summation-order noise only


def test_oracle_text_mask_is_causal():
    """Changing tokens after the end-of-text position must not change the feature (causal mask + EOT pooling)."""
    meta, ref = load_golden("text_tiny")
    tg, sd = regenerate_text(meta)
    tok = ref["tokens"].copy()
    base = O.text_encode(sd, tg, tok).numpy()
    eot = tok.argmax(-1)
    for i in range(tok.shape[0]):
        tok[i, eot[i] + 1:] = 17                        # < EOT id, so argmax is unchanged
    assert rel(O.text_encode(sd, tg, tok).numpy(), base) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_text_encode_matches_reference(lib, name):
    meta, ref = load_golden(name)
    tg, sd = regenerate_text(meta)
    g = synth.full_geometry("l14-2layer" if meta["geom"] == "ViT-L/14" else meta["geom"])   # only embed_dim matters here
    assert g["embed_dim"] == meta["embed_dim"]
    eng = lib.Engine(**dict(g, max_frames=8 if meta["geom"] in ("ViT-B/16", "ViT-L/14") else 80, max_videos=10, max_tokens=8, max_classes=64,
                            otam_lambda=0.5, device=0))
    try:
        with pytest.raises(lib.FsarError):              # not configured yet
            eng.text_encode(torch.from_numpy(ref["tokens"]))
        eng.text_configure(**tg)
        with pytest.raises(lib.FsarError):              # weights missing
            eng.text_encode(torch.from_numpy(ref["tokens"]))
        assert eng.load_clip_text_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}) == len(sd)
        n0 = eng.launch_count()
        out = eng.text_encode(torch.from_numpy(ref["tokens"])).cpu().numpy()
        assert eng.launch_count() - n0 == 2 + 7 * tg["layers"]
        # fp16 operands vs the fp32 reference: operand rounding (same bound as the frame encoder's stress weights)
        assert rel(out, ref["features"]) < 3e-3
        # against the oracle rounding where the CUDA path stores a 16-bit operand: isolates kernel bugs from rounding
        emu = O.text_encode(sd, tg, ref["tokens"], operand_dtype=eng.operand_dtype).numpy()
        # (2 layers: < 1e-3; 12 layers accumulate tanh.approx / accumulation-order differences between the two 16-bit paths)
        assert rel(out, emu) < (1.5e-3 if tg["layers"] <= 3 else 2.5e-3)
        # cosine between every pair of class embeddings (what cos_sim consumes) agrees to 1e-3
        def cosmat(f):
            f = f / np.linalg.norm(f, axis=-1, keepdims=True)
            return f @ f.T
        assert np.abs(cosmat(out) - cosmat(ref["features"])).max() < 1e-3
        # more texts than fit one workspace pass (chunking), and determinism
        many = torch.from_numpy(np.tile(ref["tokens"], (40, 1)))
        out_many = eng.text_encode(many).cpu().numpy()
        assert np.array_equal(out_many[:out.shape[0]], out)
        assert np.array_equal(out_many[-out.shape[0]:], out)
    finally:
        eng.close()


@pytest.mark.gpu
def test_gpu_text_tower_does_not_block_episodes(lib):
    """A configured but unfilled text tower must not make the episode entry points report missing weights."""
    meta, ref = load_golden("tiny_5w1s")
    from conftest import regenerate
    g, sd, tt, te, task = regenerate(meta)
    eng = lib.Engine(**dict(g, max_frames=80, max_videos=10, max_tokens=8, max_classes=64, otam_lambda=0.5, device=0))
    try:
        eng.text_configure(**synth.TEXT_GEOMETRIES["tiny"])
        eng.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        eng.set_weight("text_features_train", torch.from_numpy(tt))
        eng.set_weight("text_features_test", torch.from_numpy(te))
        dev = {k: torch.from_numpy(v).cuda() for k, v in task.items()}
        logits, _ = eng.episode_forward(dev["support_set"], dev["target_set"], dev["support_labels"],
                                        dev["real_support_labels"], meta["T"], meta["way"], n_train_classes=meta["n_train"])
        assert rel(logits.cpu().numpy(), ref["logits"]) < 3e-3
    finally:
        eng.close()


@pytest.mark.skipif(not __import__("os").path.isdir("/root/reference"), reason="reference tree not present (GPU box)")
def test_reference_tokenizer_reproduces_fixture_tokens():
    """The drop-in module's default tokenizer is the reference's own tokenize(): it must give the fixture's ids."""
    from clip_fsar_b200.register import register
    register("/root/reference")
    from models.base.few_shot import tokenize
    for name in CASES:
        meta, ref = load_golden(name)
        assert np.array_equal(tokenize(meta["prompts"]).numpy().astype(np.int32), ref["tokens"])
