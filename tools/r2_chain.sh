#!/bin/sh
# end-point sharing between neighbouring cells: bit-identity test + A/B of the chain length (run under gpurun)
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "endpoint_sharing or deterministic" > $O/r2_chain_test.log 2>&1; tail -3 $O/r2_chain_test.log
python tools/ab_scan.py chain8
FSCL_B200_NO_ENDPOINT_SHARING=1 python tools/ab_scan.py nochain
FSCL_B200_LIB=build/var/lib_chain4.so python tools/ab_scan.py chain4
FSCL_B200_LIB=build/var/lib_chain16.so python tools/ab_scan.py chain16
