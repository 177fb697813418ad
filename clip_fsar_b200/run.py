"""Launcher shim: register the sm_100a head, then hand over to the reference's unmodified runs/run.py.

    cd $CLIP_FSAR_ROOT && python -m clip_fsar_b200.run --cfg <yaml> [KEY VALUE ...]
(the reference resolves configs/pool/base.yaml relative to the CWD, utils/config.py:80-93)."""
import os
import runpy
import sys

from .register import register


def main():
    root = os.environ.get("CLIP_FSAR_ROOT", "/root/reference")
    register(root)
    os.chdir(root)
    # `python runs/run.py` puts runs/ at sys.path[0] (run.py does `from test import test`, `from train import train`);
    # runpy.run_path does not, so do it here
    sys.path.insert(0, os.path.join(root, "runs"))
    sys.argv[0] = os.path.join(root, "runs", "run.py")
    runpy.run_path(sys.argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
