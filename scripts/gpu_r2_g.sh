# Round 2, call G: tensor-core stem (tests + A/B), depthwise eligibility rule.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -x -k "stem or depthwise" 2>&1 | tail -6
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_bench_shapes.py tests/test_effnetv2.py -q -m gpu --timeout 600 -x 2>&1 | tail -5
for st in 0 1; do
  EDET_STEM_IMPL=$st timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/g_ops_d0_stem$st.json > gpurun_out/g_bench_d0_stem$st.log 2>&1
  echo "stem_impl=$st: $(tail -1 gpurun_out/g_bench_d0_stem$st.log | cut -c1-230)"
done
timeout 600 python bench.py --config v2s --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/g_ops_v2s.json > gpurun_out/g_bench_v2s.log 2>&1
echo "v2s: $(tail -1 gpurun_out/g_bench_v2s.log | cut -c1-230)"
python - <<'PY'
import json
for st in (0, 1):
  d = json.load(open('gpurun_out/g_ops_d0_stem%d.json' % st))
  print('stem_impl', st, [(r['name'], round(r['ms'] * 1e3, 1)) for r in d['ops'] if r['name'] == 'stem'])
PY
