#!/bin/sh
# A/B of the main library against every variant in build/var (run under gpurun): tools/ab_variants.sh [pytest -k expression]
cd "$(dirname "$0")/.."
if [ -n "$1" ]; then python -m pytest tests -m gpu -x -q -k "$1" 2>&1 | tail -2; fi
python tools/ab_scan.py main
for v in build/var/lib_*.so; do [ -f "$v" ] && FSCL_B200_LIB=$v python tools/ab_scan.py $(basename $v .so | sed s/lib_//); done
