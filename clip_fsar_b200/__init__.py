"""clip_fsar_b200 — the B200 (sm_100a) few-shot video inference path of CLIP-FSAR behind the reference's own
head-module boundary. Compute lives in libfsar_sm100.so (csrc/, C ABI in include/fsar.h)."""
from . import synth  # noqa: F401
from .lib import Engine, FsarError, geometry, load_library  # noqa: F401

__all__ = ["Engine", "FsarError", "geometry", "load_library", "synth"]
