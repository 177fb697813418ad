import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu")


def golden_names():
    """Episode fixtures (preproc_* belong to tests/test_preprocess.py, text_* to tests/test_text_tower.py, the full-depth
    ViT-L/14 episode has tests of its own)."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR)
                  if f.endswith(".npz") and not f.startswith(("preproc_", "text_", "vitl14_")))


def load_golden(name):
    """Fixture written by oracle/gen_golden.py from the reference's own forward."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return meta, {k: z[k] for k in z.files if k != "meta"}


def regenerate(meta):
    """Weights / text features / episode of a fixture from its seeds (clip_fsar_b200.synth, numpy PCG64)."""
    from clip_fsar_b200 import synth
    g = synth.full_geometry(meta["geom"], meta["mod_depth"])
    sd = synth.synth_state_dict(g, meta["wseed"], meta["spread"])
    tt = synth.synth_text_features(meta["n_train"], g["embed_dim"], meta["text_seeds"][0])
    te = synth.synth_text_features(meta["n_test"], g["embed_dim"], meta["text_seeds"][1])
    task = synth.synth_episode(meta["way"], meta["shot"], meta.get("qpc", 1), meta["T"], g["image_size"], meta["n_test"], meta["eseed"],
                               meta["structured"])
    if meta.get("keep_counts"):
        task = synth.ragged_support(task, meta["T"], meta["keep_counts"])
    return g, sd, tt, te, task


@pytest.fixture(scope="session")
def lib():
    from clip_fsar_b200 import build, lib as L
    build.build()
    return L
