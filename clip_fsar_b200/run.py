"""Launcher shim: register the sm_100a head, then hand over to the reference's unmodified runs/run.py.

    CLIP_FSAR_ROOT=<reference tree> python -m clip_fsar_b200.run --cfg <yaml> [--init_method tcp://127.0.0.1:PORT]
(the reference resolves configs/pool/base.yaml relative to the CWD, utils/config.py:80-93, so the shim chdirs there).

Registration happens at IMPORT of this module, not only under `__main__`: with NUM_GPUS > 1 the reference launches its
ranks with torch.multiprocessing.spawn (utils/launcher.py:29-34), and every spawned interpreter re-imports the parent's
main module before it unpickles `run(local_rank, func, init_method, cfg)`. The reference's run.py is therefore executed
inside THIS module's namespace (so that the children see `clip_fsar_b200.run` as their main module and register the head
and the import stubs too) instead of through runpy, which would make runs/run.py itself the children's main module."""
import os
import sys

from .register import register

ROOT = os.environ.get("CLIP_FSAR_ROOT", "/root/reference")


def _prepare():
    """Idempotent; runs in the launcher and in every spawned rank."""
    register(ROOT)
    runs = os.path.join(ROOT, "runs")
    # `python runs/run.py` puts runs/ at sys.path[0] (run.py does `from test import test`, `from train import train`)
    if runs not in sys.path:
        sys.path.insert(0, runs)


if os.path.isdir(os.path.join(ROOT, "models", "base")):
    _prepare()


def main():
    _prepare()
    os.chdir(ROOT)
    run_py = os.path.join(ROOT, "runs", "run.py")
    sys.argv[0] = run_py
    with open(run_py) as f:
        code = compile(f.read(), run_py, "exec")
    exec(code, sys.modules["__main__"].__dict__)      # its `if __name__ == "__main__": main()` fires here


if __name__ == "__main__":
    main()
