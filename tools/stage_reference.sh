#!/usr/bin/env bash
# Stage an UNMODIFIED copy of the reference tree under baseline/_ref/ (git-ignored, NOT gpurun-ignored), so that it
# travels to the GPU box with the repo snapshot. /root/reference itself does not exist there.
#   tools/stage_reference.sh [SRC]        (default SRC = /root/reference)
# Used by: bench.py --impl reference (times the reference's own BaseVideoModel forward on the host cores),
#          tests/test_reference_runner.py (runs/test_net_few_shot.py:test_epoch drives the sm_100a head on a B200),
#          python -m clip_fsar_b200.run (CLIP_FSAR_ROOT=baseline/_ref).
# Our two YAML files are dropped next to the config they inherit from (the only files added; nothing is edited).
set -euo pipefail
SRC="${1:-/root/reference}"
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
DST="$ROOT/baseline/_ref"
if [ ! -d "$SRC/models/base" ]; then echo "reference tree not found at $SRC" >&2; exit 1; fi
rm -rf "$DST"
mkdir -p "$DST"
# sources + configs only (no images / docs); cp -a keeps the files byte-identical
for d in configs datasets models runs sslgenerators utils; do cp -a "$SRC/$d" "$DST/$d"; done
cp -a "$SRC/LICENSE" "$DST/LICENSE" 2>/dev/null || true
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
cp "$ROOT"/configs/*_sm100.yaml "$DST/configs/projects/CLIPFSAR/kinetics100/"
( cd "$SRC" && find configs datasets models runs sslgenerators utils -type f ! -path '*/__pycache__/*' -print0 | sort -z | xargs -0 sha256sum ) > "$DST/SHA256SUMS"
echo "staged $(find "$DST" -type f | wc -l) files into $DST ($(du -sh "$DST" | cut -f1))"
