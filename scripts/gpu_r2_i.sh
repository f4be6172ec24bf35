# Round 2, call I: TMA-staged sepconv_direct (tests + A/B).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -x -k "sepconv" 2>&1 | tail -6
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_bench_shapes.py -q -m gpu --timeout 600 -x 2>&1 | tail -5
for si in 0 1; do
  EDET_SEPCONV_IMPL=$si timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/i_ops_d0_sep$si.json > gpurun_out/i_bench_d0_sep$si.log 2>&1
  echo "sepconv_impl=$si: $(tail -1 gpurun_out/i_bench_d0_sep$si.log | cut -c1-230)"
done
timeout 300 python scripts/time_segments.py d0 gpurun_out/i_segments_d0.json 2>&1 | tail -4
python - <<'PY'
import json
for si in (0, 1):
  d = json.load(open('gpurun_out/i_ops_d0_sep%d.json' % si))
  print('sepconv_impl', si, d['kinds']['sepconv_tc'])
PY
