#!/bin/bash
# quick ncu --set full of one blind-rotate variant at a small count: tools/prof_variant.sh <variant> <count>
V=$1; C=${2:-3552}
TFHE_BR_VARIANT=$V ncu --set full --clock-control none --import-source on -k regex:blind_rotate -s 2 -c 1 \
   -o gpurun_out/prof_br_v$V -f python tools/quick_bench.py 128 $C > /dev/null 2>&1
ls -la gpurun_out/prof_br_v$V.ncu-rep
