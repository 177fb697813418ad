"""CPU (needs the reference tree): the reference's own Config / YAML chain and build_model construct the sm_100a head
from the YAML files shipped in configs/ — the "YAML plumbing unchanged" half of the drop-in claim."""
import os
import shutil
import sys

import pytest

from conftest import ROOT, load_golden

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "configs")), reason="reference tree not on this box")


@pytest.mark.parametrize("yaml_name", ["CLIPFSAR_synth_5way1shot_vitb16_sm100.yaml", "CLIPFSAR_K100_1shot_vitb16_sm100.yaml"])
def test_reference_config_chain_builds_our_head(tmp_path, monkeypatch, yaml_name):
    from clip_fsar_b200.register import register
    register(REF)                                           # stubs oss2 / simplejson / decord, fills the registries
    # a scratch copy of the reference's configs/ with our YAML dropped next to the one it inherits from
    shutil.copytree(os.path.join(REF, "configs"), tmp_path / "configs")
    dst = tmp_path / "configs" / "projects" / "CLIPFSAR" / "kinetics100" / yaml_name
    shutil.copy(os.path.join(ROOT, "configs", yaml_name), dst)
    monkeypatch.chdir(tmp_path)                             # utils/config.py:86 reads ./configs/pool/base.yaml
    rel = os.path.relpath(dst, tmp_path)
    monkeypatch.setattr(sys, "argv", ["runs/run.py", "--cfg", rel])
    from utils.config import Config
    cfg = Config(load=True)
    assert cfg.VIDEO.HEAD.NAME == "CNN_OTAM_CLIPFSAR_SM100" and cfg.VIDEO.HEAD.BACKBONE_NAME == "ViT-B/16"
    assert cfg.MODEL.NAME == "BaseVideoModel" and cfg.TASK_TYPE == "few_shot_action"      # inherited through _BASE chain
    assert cfg.DATA.NUM_INPUT_FRAMES == 8 and len(cfg.TRAIN.CLASS_NAME) == 64 and len(cfg.TEST.CLASS_NAME) == 24
    assert cfg.TEST.ENABLE is True and cfg.TRAIN.ENABLE is False

    if "synth" not in yaml_name:
        import torch
        torch.save({"train": torch.randn(64, 512), "test": torch.randn(24, 512)}, tmp_path / "text_features_k100_vitb16.pt")
    cfg.NUM_GPUS = 0                                        # no GPU in this container: skip .cuda() / DDP in build_model
    from models.base.builder import build_model
    model, _ = build_model(cfg)
    from clip_fsar_b200.head import CNN_OTAM_CLIPFSAR_SM100
    assert isinstance(model.head, CNN_OTAM_CLIPFSAR_SM100)
    assert tuple(model.head.text_features_train.shape) == (64, 512) and tuple(model.head.text_features_test.shape) == (24, 512)
    meta, _ = load_golden("vitb16_5w1s")
    assert sorted(model.state_dict().keys()) == sorted("head." + k for k in meta["state_dict_keys"])

    if "synth" in yaml_name:                                # the reference's loader builds our synthetic dataset
        from datasets.base.builder import build_dataset
        ds = build_dataset(cfg.TEST.DATASET, cfg, "test")
        assert len(ds) == 200 and ds[0]["support_set"].shape == (40, 3, 224, 224)
