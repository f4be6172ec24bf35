set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -k "pointwise" --timeout 120 2>&1 | tail -12
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/ops_r1h.json > gpurun_out/bench8.log 2>&1; tail -1 gpurun_out/bench8.log | cut -c1-400
