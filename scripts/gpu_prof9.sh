#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m automl_b200.build > gpurun_out/build9.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:mbconv_front -s 4 -c 1 -f -o gpurun_out/mbf_b1 python scripts/bench_mbconv.py 32 b1 > gpurun_out/ncu_b1.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:mbconv_front -s 4 -c 1 -f -o gpurun_out/mbf_b4 python scripts/bench_mbconv.py 32 b4 > gpurun_out/ncu_b4.log 2>&1
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
