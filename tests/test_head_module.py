"""CPU: the drop-in nn.Module mirrors the reference head's interface (names, ctor, registry) and never computes on CPU."""
import os
import types

import numpy as np
import pytest
import torch

from conftest import load_golden

NS = types.SimpleNamespace


def make_cfg(name="CNN_OTAM_CLIPFSAR_SM100", backbone="ViT-B/16", T=8, **train):
    return NS(TRAIN=NS(CLASS_NAME=["c%d" % i for i in range(64)], WAY=5, SHOT=1, BATCH_SIZE=1, **train),
              TEST=NS(CLASS_NAME=["t%d" % i for i in range(24)]), DATA=NS(NUM_INPUT_FRAMES=T),
              VIDEO=NS(HEAD=NS(NAME=name, BACKBONE_NAME=backbone, SYNTHETIC_TEXT=True), BACKBONE=NS(META_ARCH="Identity")),
              BN=NS(FREEZE=False))


def test_state_dict_keys_and_shapes_equal_reference_head():
    from clip_fsar_b200.head import CNN_OTAM_CLIPFSAR_SM100
    from clip_fsar_b200 import synth
    meta, _ = load_golden("vitb16_5w1s")
    head = CNN_OTAM_CLIPFSAR_SM100(make_cfg())
    sd = head.state_dict()
    assert sorted(sd.keys()) == meta["state_dict_keys"]
    shapes = synth.state_dict_shapes(head.geometry)
    assert {k: tuple(v.shape) for k, v in sd.items()} == shapes
    assert sum(p.numel() for p in head.parameters()) == 89342465


def test_transformer_depth_adds_layers():
    from clip_fsar_b200.head import CNN_OTAM_CLIPFSAR_SM100
    meta, _ = load_golden("tiny_5w1s_depth2")
    head = CNN_OTAM_CLIPFSAR_SM100(make_cfg(backbone="tiny", TRANSFORMER_DEPTH=2))
    assert sorted(head.state_dict().keys()) == meta["state_dict_keys"]


def test_load_state_dict_round_trip_marks_engine_dirty():
    from clip_fsar_b200.head import CNN_OTAM_CLIPFSAR_SM100
    from clip_fsar_b200 import synth
    head = CNN_OTAM_CLIPFSAR_SM100(make_cfg(backbone="tiny"))
    sd = {k: torch.from_numpy(v) for k, v in synth.synth_state_dict(head.geometry, 3).items()}
    head._pushed_versions = ("stale",)
    res = head.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert head._pushed_versions is None
    assert torch.equal(head.state_dict()["backbone.proj"], sd["backbone.proj"])


def test_unsupported_branches_raise():
    from clip_fsar_b200.head import CNN_OTAM_CLIPFSAR_SM100
    # the text branches are part of the path: EVAL_TEXT wins over COMBINE like the reference's if / elif
    assert CNN_OTAM_CLIPFSAR_SM100(make_cfg(backbone="tiny", EVAL_TEXT=True, COMBINE=True)).text_mode == 1
    h = CNN_OTAM_CLIPFSAR_SM100(make_cfg(backbone="tiny", COMBINE=True, TEXT_COFF=0.5))
    assert h.text_mode == 2 and h.text_coff == 0.5
    assert CNN_OTAM_CLIPFSAR_SM100(make_cfg(backbone="tiny", COMBINE=True)).text_coff == 0.9
    with pytest.raises(ValueError):
        CNN_OTAM_CLIPFSAR_SM100(make_cfg(backbone="RN50"))
    cfg = make_cfg(backbone="tiny")
    cfg.VIDEO.HEAD.SYNTHETIC_TEXT = False
    with pytest.raises(ValueError):
        CNN_OTAM_CLIPFSAR_SM100(cfg)


def test_forward_never_computes_on_cpu():
    from clip_fsar_b200.head import CNN_OTAM_CLIPFSAR_SM100
    from clip_fsar_b200.lib import FsarError
    head = CNN_OTAM_CLIPFSAR_SM100(make_cfg(backbone="tiny"))
    task = {"support_set": torch.zeros(40, 3, 32, 32), "target_set": torch.zeros(40, 3, 32, 32),
            "support_labels": torch.arange(5.0), "real_support_labels": torch.arange(5.0)}
    with pytest.raises(NotImplementedError):
        head.train()(task)          # training mode is not part of the path
    with pytest.raises(FsarError):
        head.eval()(task)           # CPU tensors: loud failure, no fallback


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree not on this box")
def test_registers_into_reference_registry_and_builds_through_it():
    from clip_fsar_b200.register import register
    heads, datasets = register()
    assert heads.get("CNN_OTAM_CLIPFSAR_SM100") is not None
    assert heads.get("CNN_OTAM_CLIPFSAR") is not None          # the reference head is untouched
    assert datasets.get("Synth_few_shot") is not None
    register()                                                  # idempotent (Registry asserts on duplicates)
    from models.base.models import BaseVideoModel
    model = BaseVideoModel(make_cfg(backbone="tiny")).eval()
    keys = sorted(model.state_dict().keys())
    meta, _ = load_golden("tiny_5w1s")
    assert keys == sorted("head." + k for k in meta["state_dict_keys"])   # what utils/checkpoint.py:329 looks up


@pytest.mark.skipif(not os.path.isdir("/root/reference/models"), reason="reference tree not on this box")
def test_synthetic_dataset_matches_reference_task_dict():
    from clip_fsar_b200.register import register
    _, datasets = register()
    cfg = make_cfg(backbone="tiny")
    cfg.TRAIN.NUM_TEST_TASKS = 4
    cfg.DATA.TEST_CROP_SIZE = 32
    ds = datasets.get("Synth_few_shot")(cfg, "test")
    item = ds[1]
    assert len(ds) == 4
    assert set(item) == {"support_set", "support_labels", "target_set", "target_labels", "real_support_labels",
                         "real_target_labels", "batch_class_list"}     # ssv2_few_shot.py:275-285
    assert item["support_set"].shape == (40, 3, 32, 32) and item["support_set"].dtype == torch.float32
    assert sorted(item["support_labels"].tolist()) == [0.0, 1.0, 2.0, 3.0, 4.0]
    assert torch.equal(ds[1]["target_set"], item["target_set"])        # seeded by index


def test_clip_checkpoint_fills_backbone_and_schedules_text_tower(tmp_path):
    """VIDEO.HEAD.CLIP_CHECKPOINT: 'visual.*' -> backbone.*, the rest -> the library's text tower at the first forward
    (few_shot.py:2706-2728 does load(...) + encode_text in the constructor)."""
    from clip_fsar_b200 import synth
    from clip_fsar_b200.head import CNN_OTAM_CLIPFSAR_SM100
    g = synth.full_geometry("tiny")
    vis = {k[len("backbone."):]: torch.from_numpy(v) for k, v in synth.synth_state_dict(g, 5).items() if k.startswith("backbone.")}
    tg = synth.TEXT_GEOMETRIES["tiny"]
    txt = {k: torch.from_numpy(v) for k, v in synth.synth_text_state_dict(tg, g["embed_dim"], 3).items()}
    ckpt = {("visual." + k): v for k, v in vis.items()}
    ckpt.update(txt)
    ckpt["logit_scale"] = torch.tensor(4.6)
    path = os.path.join(tmp_path, "clip_tiny.pt")
    torch.save(ckpt, path)
    cfg = make_cfg(backbone="tiny")
    cfg.VIDEO.HEAD.SYNTHETIC_TEXT = False
    cfg.VIDEO.HEAD.CLIP_CHECKPOINT = path
    cfg.TEST.PROMPT = "a video of {}"
    prompts = []

    def tokenizer(ps):
        prompts.append(list(ps))
        return torch.zeros(len(ps), 77, dtype=torch.int32)

    head = CNN_OTAM_CLIPFSAR_SM100(cfg, tokenizer=tokenizer)
    assert torch.equal(head.backbone.proj, vis["proj"])
    assert torch.equal(head.backbone.transformer.resblocks[1].mlp.c_fc.weight, vis["transformer.resblocks.1.mlp.c_fc.weight"])
    assert head._text_pending and head._text_geometry == tg
    assert prompts[0][0] == "a video of c0" and prompts[1][0] == "a video of t0" and len(prompts[1]) == 24
    assert "logit_scale" not in head._text_state and "token_embedding.weight" in head._text_state
    # explicit features cancel the pending text-tower run
    head.set_text_features(torch.zeros(64, g["embed_dim"]), torch.zeros(24, g["embed_dim"]))
    assert not head._text_pending
