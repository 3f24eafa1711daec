#!/bin/sh
# build a debug variant of the CUDA library into build/var/: tools/build_variant.sh hist|timing [extra nvcc flags]
set -e
cd "$(dirname "$0")/.."
mkdir -p build/var
case "$1" in
  hist) DEF=-DFSCL_HIST ;;
  timing) DEF=-DFSCL_TIMING ;;
  *) DEF="" ;;
esac
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $DEF "$@" \
     -o build/var/lib_$name.so fscl_b200/csrc/fscl_kernels.cu fscl_b200/csrc/fscl_host.cu fscl_b200/csrc/fscl_tables.cu -ldl
echo build/var/lib_$name.so
