#!/bin/bash
# Stage the UNMODIFIED reference checkout for the GPU box (which has no /root/reference): one tarball under baseline/_ref/
# (git-ignored, travels with the gpurun snapshot).  Only tests/test_gpu_reference_scripts.py and tools/run_reference_script.py
# read it; nothing of the product does.  Usage: tools/stage_reference.sh [/path/to/FORGE]
set -e
REF=${1:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/baseline/_ref"
tar -C "$REF" --exclude=.git --exclude='__pycache__' -czf "$ROOT/baseline/_ref/forge_reference.tar.gz" .
ls -la "$ROOT/baseline/_ref/forge_reference.tar.gz"
