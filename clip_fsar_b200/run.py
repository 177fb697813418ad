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
    sys.argv[0] = os.path.join(root, "runs", "run.py")
    runpy.run_path(sys.argv[0], run_name="__main__")


if __name__ == "__main__":
    main()
