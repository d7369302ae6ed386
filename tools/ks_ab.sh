#!/bin/bash
# A/B of the tcgen05 key switch (default) against the mma.sync one + blind-rotate variant 9; parity first.
set -x
mkdir -p gpurun_out
timeout 300 python tools/sanitize.py 2>&1 | tail -4
TFHE_KS_VARIANT=mma timeout 300 python tools/sanitize.py 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -4
timeout 300 python tools/quick_bench.py 128 601,16384,65536 2>&1 | grep "rep=2"
TFHE_KS_VARIANT=mma timeout 300 python tools/quick_bench.py 128 65536 2>&1 | grep "rep=2"
TFHE_BR_VARIANT=9 timeout 300 python tools/quick_bench.py 128 65536 2>&1 | grep "rep=2"
timeout 300 python tools/quick_bench.py 80 65536 2>&1 | grep "rep=2"
timeout 300 python tools/quick_bench.py 110 65536 2>&1 | grep "rep=2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ks_umma -s 1 -c 1 \
   -o gpurun_out/prof_ks_umma -f python tools/quick_bench.py 128 65536 > /dev/null 2>&1
ls -la gpurun_out/
