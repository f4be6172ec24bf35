# Round 2, call E: dynamic tile scheduling in pointwise_tc / dw_tile: tests + bench + segments.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -x -k "pointwise or depthwise or se_fc" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_bench_shapes.py tests/test_gpu_serving.py -q -m gpu --timeout 600 -x 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/e_ops_d0.json > gpurun_out/e_bench_d0.log 2>&1
echo "d0: $(tail -1 gpurun_out/e_bench_d0.log | cut -c1-230)"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/e_bench_d0_b.log 2>&1
echo "d0 again: $(tail -1 gpurun_out/e_bench_d0_b.log | cut -c1-230)"
timeout 300 python scripts/time_segments.py d0 gpurun_out/e_segments_d0.json 2>&1 | tail -4
for c in d4 d7x; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/e_bench_$c.log 2>&1
  echo "$c: $(tail -1 gpurun_out/e_bench_$c.log | cut -c1-230)"
done
