"""Register the sm_100a head (and a synthetic episode dataset) into the reference's own registries.

The reference builds its head with `HEAD_REGISTRY.get(cfg.VIDEO.HEAD.NAME)(cfg=cfg)` (models/base/models.py:40) and
its dataset with `DATASET_REGISTRY.get(name.capitalize())(cfg, split)` (datasets/base/builder.py:113-124); both key
by class `__name__` and assert on duplicates (utils/registry.py:27-49). Importing this module from a launcher
before `runs/run.py` starts is all that is needed; `runs/*.py` and the YAML plumbing stay byte-identical:

    python -m clip_fsar_b200.run --cfg configs/projects/CLIPFSAR/kinetics100/CLIPFSAR_K100_1shot_v1.yaml
    # with  VIDEO.HEAD.NAME: CNN_OTAM_CLIPFSAR_SM100, VIDEO.HEAD.BACKBONE_NAME: "ViT-B/16" in the YAML
"""
import os
import sys
import types

import numpy as np
import torch

from . import synth
from .head import CNN_OTAM_CLIPFSAR_SM100

_DONE = {}


def _stub(name, **kw):
    if name in sys.modules:
        return
    try:
        __import__(name)
    except ImportError:
        m = types.ModuleType(name)
        m.__dict__.update(kw)
        sys.modules[name] = m


def _simplejson_dumps(obj, use_decimal=False, **kw):
    """simplejson.dumps as utils/logging.py:86 calls it (sort_keys=True, use_decimal=True on a dict whose floats were
    wrapped in decimal.Decimal): the stdlib encoder with Decimals written as plain numbers."""
    import decimal
    import json
    return json.dumps(obj, default=lambda o: float(o) if isinstance(o, decimal.Decimal) else str(o), **kw)


def stub_optional_dependencies():
    """Modules the reference imports at module scope but never needs on the few-shot inference path; absent in
    this image. Real installs are used when present."""
    _stub("ipdb", set_trace=lambda *a, **k: None)                      # few_shot.py:15
    _stub("ftfy", fix_text=lambda s: s)                                # few_shot.py:30
    _stub("oss2")                                                      # test_net_few_shot.py:12
    _stub("simplejson", dumps=_simplejson_dumps, loads=__import__("json").loads)          # utils/logging.py:16, 86
    if "decord" not in sys.modules:
        try:
            import decord  # noqa: F401
        except ImportError:
            d = types.ModuleType("decord")
            d.VideoReader = object
            d.cpu = lambda *a, **k: None
            d.gpu = lambda *a, **k: None
            d.bridge = types.SimpleNamespace(set_bridge=lambda *a, **k: None)
            sys.modules["decord"] = d


class Synth_few_shot(torch.utils.data.Dataset):
    """Synthetic stand-in for datasets/base/ssv2_few_shot.py:Ssv2_few_shot: same task-dict keys, dtypes and
    shapes (190-285), seeded by the global episode index so runs are reproducible at any world size.
    Select with TEST.DATASET: synth_few_shot (the builder capitalises the name)."""

    def __init__(self, cfg, split):
        self.cfg = cfg
        self.split = split
        self.way = int(getattr(cfg.TRAIN, "WAY_TEST", None) or getattr(cfg.TRAIN, "WAY", 5))
        self.shot = int(getattr(cfg.TRAIN, "SHOT_TEST", None) or getattr(cfg.TRAIN, "SHOT", 1))
        self.queries = int(getattr(cfg.TRAIN, "QUERY_PER_CLASS_TEST", None) or getattr(cfg.TRAIN, "QUERY_PER_CLASS", 1))
        self.frames = int(cfg.DATA.NUM_INPUT_FRAMES)
        self.size = int(getattr(cfg.DATA, "TEST_CROP_SIZE", 224))
        self.n_cls = len(getattr(cfg.TEST, "CLASS_NAME", []) or [])
        if self.n_cls < self.way:
            # real_support_labels index the rows of text_features_test (few_shot.py:2946), one per TEST.CLASS_NAME entry
            raise ValueError("TEST.CLASS_NAME lists %d classes but the episodes are %d-way: real labels would index "
                             "past text_features_test" % (self.n_cls, self.way))
        self.length = int(getattr(cfg.TRAIN, "NUM_TEST_TASKS", 100))
        # FSAR_SYNTH_POOL=n: cycle n cached episodes (per loader worker) instead of generating every index afresh, so a
        # throughput run through the reference runner measures the runner, not numpy's random generator
        self.pool = int(os.environ.get("FSAR_SYNTH_POOL", "0") or 0)
        self._cache = {}

    def __len__(self):
        return self.length

    def __getitem__(self, index):
        index = int(index)
        if self.pool > 0:
            index %= self.pool
            if index in self._cache:
                return self._cache[index]
        ep = synth.synth_episode(self.way, self.shot, self.queries, self.frames, self.size, self.n_cls, 1000 + index)
        item = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in ep.items()}
        if self.pool > 0:
            self._cache[index] = item
        return item


def register(reference_root=None):
    """Put CNN_OTAM_CLIPFSAR_SM100 into HEAD_REGISTRY and Synth_few_shot into DATASET_REGISTRY. Idempotent.
    Returns the two registries. Raises ImportError when the reference tree is not importable."""
    root = reference_root or os.environ.get("CLIP_FSAR_ROOT", "/root/reference")
    if root in _DONE:
        return _DONE[root]
    if not os.path.isdir(os.path.join(root, "models", "base")):
        raise ImportError("reference tree not found at %r (set CLIP_FSAR_ROOT)" % root)
    if root not in sys.path:
        sys.path.insert(0, root)
    stub_optional_dependencies()
    from models.base.base_blocks import HEAD_REGISTRY
    if HEAD_REGISTRY.get(CNN_OTAM_CLIPFSAR_SM100.__name__) is None:
        HEAD_REGISTRY.register()(CNN_OTAM_CLIPFSAR_SM100)
    from datasets.base.builder import DATASET_REGISTRY
    if DATASET_REGISTRY.get(Synth_few_shot.__name__) is None:
        DATASET_REGISTRY.register()(Synth_few_shot)
    _DONE[root] = (HEAD_REGISTRY, DATASET_REGISTRY)
    return _DONE[root]
