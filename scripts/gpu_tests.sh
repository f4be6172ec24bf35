# Build + the whole GPU test suite + one bench line with the per-op profile.
# usage (from the repo root, on the dev container):  gpurun --timeout 900 -- 'bash scripts/gpu_tests.sh'
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -x 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/ops.json > gpurun_out/bench.log 2>&1
tail -1 gpurun_out/bench.log | cut -c1-400
