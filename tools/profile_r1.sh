#!/bin/bash
# Round profile recipe (run under gpurun): bench line, ncu launch list, full captures of
# the two hot kernels.  Numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/clocks_before.csv
python bench.py --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 3000 gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2>> gpurun_out/bench.err
ncu --set full --clock-control none --import-source on -k regex:blind_rotate -s 1 -c 1 -o gpurun_out/prof_br -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>> gpurun_out/bench.err
ncu --set full --clock-control none --import-source on -k regex:ks_mma -s 1 -c 1 -o gpurun_out/prof_ks -f \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>> gpurun_out/bench.err
tail -5 gpurun_out/bench.err
ls -la gpurun_out
