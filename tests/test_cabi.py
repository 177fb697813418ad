"""CPU: the C-ABI shared library builds, loads and exports exactly what include/fsar.h declares; it refuses to
run without an sm_100 device (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "fsar.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fsar_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree(lib):
    assert header_symbols() == sorted(lib.SYMBOLS)


def test_library_exports_every_declared_symbol(lib):
    so = ctypes.CDLL(lib.LIB_PATH)
    for name in header_symbols():
        assert hasattr(so, name), "libfsar_sm100.so does not export %s" % name
    assert lib.load_library().fsar_version() == 100
    assert lib.load_library().fsar_class_name(4) == b"attention"
    assert lib.load_library().fsar_operand_dtype() in (0, 1)


def test_config_struct_layout(lib):
    # 17 x 4-byte fields, in header order
    assert ctypes.sizeof(lib.FsarConfig) == 68
    assert ctypes.sizeof(lib.FsarEpisode) == 4 * 8 + 8 * 4
    assert ctypes.sizeof(lib.FsarProfile) == 13 * 8 * 4 and lib.FSAR_PROF_CLASSES == 13
    assert ctypes.sizeof(lib.FsarTextConfig) == 20


def test_library_has_no_driver_link_dependency(lib):
    # cudart is linked statically and the driver is reached through cudaGetDriverEntryPoint, so the library loads on
    # a box without libcuda (this container) and fails at fsar_create, loudly, instead of at dlopen
    import subprocess
    out = subprocess.run(["ldd", lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libcuda.so" not in out and "libcudart" not in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu(lib):
    with pytest.raises(lib.FsarError) as e:
        lib.Engine(**lib.geometry("ViT-B/16"))
    assert e.value.code == -2 and "no CPU path" in str(e.value)


def test_create_rejects_bad_geometry(lib):
    g = lib.geometry("ViT-B/16")
    g["width"] = 700
    with pytest.raises(lib.FsarError) as e:
        lib.Engine(**g)
    assert e.value.code == -1


def test_missing_library_is_an_error(lib, tmp_path):
    with pytest.raises(FileNotFoundError):
        lib.load_library(str(tmp_path / "libfsar_sm100.so"))


def test_geometry_table(lib):
    b16 = lib.geometry("ViT-B/16", num_frames=8)
    assert (b16["width"], b16["layers"], b16["heads"], b16["embed_dim"], b16["mod_dim_head"]) == (768, 12, 12, 512, 64)
    l14 = lib.geometry("ViT-L/14", num_frames=16)
    assert (l14["width"], l14["layers"], l14["heads"], l14["embed_dim"], l14["patch_size"]) == (1024, 24, 16, 768, 14)
    with pytest.raises(ValueError):
        lib.geometry("RN50")
