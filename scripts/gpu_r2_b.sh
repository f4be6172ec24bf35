# Round 2, call B: GPU suite after the per-device state / pipelined serving changes, and a bench
# line for every BASELINE configuration.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -15 | tee gpurun_out/b_tests.log
for c in d0 d4 d7x v2s; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/b_ops_$c.json > gpurun_out/b_bench_$c.log 2>&1
  tail -1 gpurun_out/b_bench_$c.log | cut -c1-700
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b_ref.log 2>&1; tail -1 gpurun_out/b_ref.log | cut -c1-400
