#!/bin/bash
# Stages the seven reference modules the binding imports into the git-ignored baseline/_ref/ so that
# tests/test_gpu_integration.py can run the UNMODIFIED reference driver on a GPU box (where /root/reference does not
# exist).  Nothing is copied into the tracked tree.
set -e
src=${NNGP_REFERENCE_DIR:-/root/reference}
dst="$(dirname "$0")/../baseline/_ref"
mkdir -p "$dst"
for m in utils systems configs RK solver models parareal; do cp "$src/$m.py" "$dst/$m.py"; done
echo "staged $(ls "$dst" | wc -l) modules in $dst"
