"""GPU: the registered nn.Module (the drop-in boundary) drives the same C-ABI path and agrees with the fixtures."""
import types

import numpy as np
import pytest
import torch

from conftest import load_golden, regenerate
from test_head_module import make_cfg

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,flags", [("tiny_5w1s", {}), ("tiny_5w5s_merge", {"MERGE_BEFORE": True}),
                                        ("tiny_3w2s_T32_single", {"SINGLE_DIRECT": True}),
                                        ("tiny_5w1s_depth2", {"TRANSFORMER_DEPTH": 2}),
                                        ("tiny_5w5s_evaltext", {"EVAL_TEXT": True}),
                                        ("tiny_5w5s_combine_coff05_merge", {"COMBINE": True, "TEXT_COFF": 0.5, "MERGE_BEFORE": True}),
                                        # default-init weights: north_star's 1e-3 bound through the registered module
                                        ("tiny_5w1s_default_init", {}), ("tiny_5w5s_merge_di", {"MERGE_BEFORE": True}),
                                        ("tiny_3w2s_T32_single_di", {"SINGLE_DIRECT": True}),
                                        ("tiny_5w1s_depth2_di", {"TRANSFORMER_DEPTH": 2}),
                                        ("tiny_5w1s_combine_di", {"COMBINE": True})])
def test_module_forward_matches_reference_fixture(name, flags):
    from clip_fsar_b200.head import CNN_OTAM_CLIPFSAR_SM100
    meta, ref = load_golden(name)
    g, sd, tt, te, task = regenerate(meta)
    head = CNN_OTAM_CLIPFSAR_SM100(make_cfg(backbone=meta["geom"], T=meta["T"], **flags), torch.from_numpy(tt),
                                   torch.from_numpy(te)).cuda().eval()
    head.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    dev = {k: torch.from_numpy(v).cuda() for k, v in task.items()}
    with torch.no_grad():
        out = head(dev)
    assert set(out) == {"logits", "class_logits"}
    assert out["logits"].is_cuda and out["logits"].shape == ref["logits"].shape
    err = (out["logits"].cpu() - torch.from_numpy(ref["logits"])).abs().max() / abs(ref["logits"]).max()
    assert float(err) < (3e-3 if meta["spread"] else 1e-3)
    if not (flags.get("EVAL_TEXT") or flags.get("COMBINE")):
        top2 = np.sort(ref["logits"], axis=1)[:, -2:]
        sure = (top2[:, 1] - top2[:, 0]) > 2e-3 * abs(ref["logits"]).max()
        assert (out["logits"].cpu().numpy().argmax(1)[sure] == ref["logits"].argmax(1)[sure]).all()
    else:
        assert out["class_logits"] is None          # few_shot.py:2852 / 2930
        return
    # way falls back to torch.unique when batch_class_list is absent (few_shot.py:2965)
    dev.pop("batch_class_list")
    with torch.no_grad():
        again = head(dev)
    assert torch.equal(again["logits"], out["logits"])
    # parameter updates reach the engine (version counters), e.g. after loading another checkpoint
    with torch.no_grad():
        head.scale.mul_(2.0)
        scaled = head(dev)
    assert torch.allclose(scaled["class_logits"], 2.0 * out["class_logits"], rtol=1e-5, atol=1e-7)
    loss = head.loss({"target_labels": dev["target_labels"]}, out)
    assert torch.isfinite(loss)


def test_module_computes_text_features_with_the_text_tower():
    """few_shot.py:2714-2728 through the drop-in module: class prompts -> (reference tokenizer's ids, from the fixture)
    -> fsar_text_encode at the first forward. text_features_{train,test} must equal the reference's encode_text."""
    from clip_fsar_b200 import synth
    from clip_fsar_b200.head import CNN_OTAM_CLIPFSAR_SM100
    tmeta, tref = load_golden("text_tiny")
    names = [p[len("a photo of "):] for p in tmeta["prompts"]]
    meta, ref = load_golden("tiny_5w1s")
    g, sd, tt, te, task = regenerate(meta)
    cfg = make_cfg(backbone="tiny", T=meta["T"])
    cfg.TRAIN.CLASS_NAME, cfg.TEST.CLASS_NAME = names, names
    seen = []

    def tokenizer(prompts):
        seen.append(list(prompts))
        return torch.from_numpy(tref["tokens"])

    head = CNN_OTAM_CLIPFSAR_SM100(cfg).cuda().eval()
    tg = synth.TEXT_GEOMETRIES["tiny"]
    head.set_clip_text({k: torch.from_numpy(v) for k, v in synth.synth_text_state_dict(tg, 128, tmeta["wseed"]).items()},
                       tokenizer=tokenizer)
    assert seen == [tmeta["prompts"], tmeta["prompts"]] and head._text_geometry == tg
    head.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    dev = {k: torch.from_numpy(v).cuda() for k, v in task.items()}
    dev["real_support_labels"] = dev["real_support_labels"] % len(names)
    with torch.no_grad():
        out = head(dev)
    assert torch.isfinite(out["logits"]).all() and out["class_logits"].shape[1] == len(names)
    for f in (head.text_features_train, head.text_features_test):
        err = (f - torch.from_numpy(tref["features"])).abs().max() / abs(tref["features"]).max()
        assert float(err) < 3e-3
    # a second forward re-uses the features (no second encode): launch count of one episode only
    n0 = head.engine().launch_count()
    with torch.no_grad():
        again = head(dev)
    assert torch.equal(again["logits"], out["logits"])
    assert head.engine().launch_count() - n0 < 60
