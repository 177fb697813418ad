"""Build libfsar_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "fsar.cu")
OUT = os.path.join(HERE, "libfsar_sm100.so")
DEPS = sorted(os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))
              if f.endswith((".cu", ".cuh"))) + [os.path.join(HERE, "..", "include", "fsar.h")]


OUT_BF16 = os.path.join(HERE, "libfsar_sm100_bf16.so")
OUT_PROBES = os.path.join(HERE, "libfsar_sm100_probes.so")   # -DFSAR_PROBES: bottleneck probes for tools/gemm_probe.py


def nvcc_cmd(extra=(), out=OUT):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    return [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
            "-Xcompiler", "-fPIC", "-cudart", "static", *extra, "-o", out, SRC]


def up_to_date(out=OUT):
    if not os.path.exists(out):
        return False
    t = os.path.getmtime(out)
    return all(os.path.getmtime(d) <= t for d in DEPS)


def build(force=False, verbose=False, bf16=False, probes=False):
    """fp16 operands (default, libfsar_sm100.so) or bf16 operands (libfsar_sm100_bf16.so, select it with
    FSAR_LIB_PATH). Same sources, -DFSAR_BF16 switches the operand type of every 16-bit buffer."""
    out = OUT_PROBES if probes else (OUT_BF16 if bf16 else OUT)
    if not force and up_to_date(out):
        return out
    extra = ["-Xptxas", "-v"] if verbose else []
    if bf16:
        extra.append("-DFSAR_BF16")
    if probes:
        extra.append("-DFSAR_PROBES")
    r = subprocess.run(nvcc_cmd(extra, out), capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building %s" % out)
    if verbose:
        sys.stderr.write(r.stderr)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, bf16="--bf16" in sys.argv, probes="--probes" in sys.argv))
