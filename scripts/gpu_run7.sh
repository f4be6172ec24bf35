set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 2>&1 | tail -30 > gpurun_out/t_all.log; tail -8 gpurun_out/t_all.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/ops_r1g.json > gpurun_out/bench7.log 2>&1; tail -1 gpurun_out/bench7.log | cut -c1-1500
