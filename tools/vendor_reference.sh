#!/usr/bin/env bash
# Copy the UNMODIFIED reference checkout into baseline/_ref/ (git-ignored, NOT gpurun-ignored, so it
# travels to the GPU box where /root/reference does not exist).  The reference is plain Python with
# no setup.py / pyproject.toml, so "installing" it is copying its packages; nothing is built and
# nothing is edited.  Images, notebooks and dataset archives are left out (not code, 0.6 MB+).
#   tools/vendor_reference.sh [/root/reference]
set -euo pipefail
SRC="${1:-/root/reference}"
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
DST="$ROOT/baseline/_ref"
if [ ! -d "$SRC/layers/flows" ]; then
    echo "vendor_reference: $SRC is not a CategoricalNF checkout" >&2
    exit 1
fi
rm -rf "$DST"
mkdir -p "$DST"
(cd "$SRC" && find general layers experiments -type f \( -name '*.py' -o -name '*.md' \) -print0 \
    | cpio --null -pdm --quiet "$DST" 2>/dev/null) || \
(cd "$SRC" && find general layers experiments -type f -name '*.py' | while read -r f; do
    mkdir -p "$DST/$(dirname "$f")"; cp "$f" "$DST/$f"; done)
cp "$SRC/LICENSE" "$SRC/README.md" "$DST/" 2>/dev/null || true
# provenance: file list + sha256 of every vendored source, so a reader can see it is the unmodified tree
(cd "$DST" && find . -type f -name '*.py' | sort | xargs sha256sum) > "$DST/SHA256SUMS"
echo "vendored $(find "$DST" -name '*.py' | wc -l) python files from $SRC into $DST"
