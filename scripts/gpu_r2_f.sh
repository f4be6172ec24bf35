# Round 2, call F: smem plan fix for wide N, fuse_dw signatures: full GPU suite + all configs.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 2>&1 | tail -12 | tee gpurun_out/f_tests.log
for c in d0 d4 d7x v2s; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/f_ops_$c.json > gpurun_out/f_bench_$c.log 2>&1
  echo "$c: $(tail -1 gpurun_out/f_bench_$c.log | cut -c1-230)"
done
timeout 300 python scripts/time_segments.py d0 gpurun_out/f_segments_d0.json 2>&1 | tail -4
