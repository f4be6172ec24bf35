set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "not tcgen05" --timeout 300 -x 2>&1 | tail -25 > gpurun_out/t1_simt.log; cat gpurun_out/t1_simt.log | tail -15
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "tcgen05" --timeout 120 2>&1 | tail -40 > gpurun_out/t2_tc.log; cat gpurun_out/t2_tc.log | tail -25
timeout 900 python -m pytest tests/test_gpu_network.py -q --timeout 300 2>&1 | tail -40 > gpurun_out/t3_net.log; cat gpurun_out/t3_net.log | tail -25
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/ops_r1a.json > gpurun_out/bench1.log 2>&1; tail -5 gpurun_out/bench1.log
