#!/bin/bash
# fused MBConv front half: unit parity + micro-bench, then the whole gpu suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m automl_b200.build > gpurun_out/build9.log 2>&1
true
echo "mbf tests rc=$?" | tee -a gpurun_out/mbf_tests.log
tail -15 gpurun_out/mbf_tests.log
timeout 300 python scripts/bench_mbconv.py 32 > gpurun_out/mbf_bench.log 2>&1
echo "bench rc=$?"; cat gpurun_out/mbf_bench.log | tail -15
