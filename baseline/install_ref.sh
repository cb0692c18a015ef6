#!/usr/bin/env bash
# Copies the UNMODIFIED reference (Wu0409/DuPL) into baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the
# GPU box with the snapshot).  The reference is a script tree without setup.py / pyproject, so `pip install --target` has
# nothing to install: a verbatim copy of the source tree is the install.  Paper figures, the authors' training logs and the
# COCO label tables (26 MB, unused by the VOC step) stay behind.
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
DST="$HERE/_ref"
if [ ! -d "$SRC/model" ]; then
  echo "install_ref: no reference tree at $SRC" >&2
  exit 1
fi
rm -rf "$DST"
mkdir -p "$DST"
( cd "$SRC" && tar cf - --exclude=paper --exclude=logs --exclude=.git --exclude=__pycache__ \
      --exclude='datasets/coco/*.npy' --exclude='datasets/coco/*.txt' . ) | ( cd "$DST" && tar xf - )
( cd "$SRC" && find . -type f -name '*.py' -print0 | sort -z | xargs -0 sha1sum ) > "$DST/.SOURCE_SHA1"
echo "install_ref: $(find "$DST" -type f | wc -l) files, $(du -sh "$DST" | cut -f1) -> $DST"
