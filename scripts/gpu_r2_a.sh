# Round 2, call A: the whole GPU suite (old + new bench-shape / TF-semantics tests) and a bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tail -1
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 --durations=8 2>&1 | tail -40 | tee gpurun_out/a_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/a_ops.json > gpurun_out/a_bench.log 2>&1
tail -1 gpurun_out/a_bench.log | cut -c1-600
