# Round 2, call C: the TMA-tiled depthwise kernel: kernel tests, A/B against the register kernel,
# network tests, bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tf_semantics.py -q -m gpu --timeout 300 -x -k "depthwise" 2>&1 | tail -15 | tee gpurun_out/c_tests_dw.log
timeout 300 python scripts/ab_depthwise.py gpurun_out/c_ab_depthwise.json 2>&1 | tail -20
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_bench_shapes.py tests/test_effnetv2.py -q -m gpu --timeout 600 -x 2>&1 | tail -8 | tee gpurun_out/c_tests_net.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/c_ops_d0.json > gpurun_out/c_bench_d0.log 2>&1
tail -1 gpurun_out/c_bench_d0.log | cut -c1-400
