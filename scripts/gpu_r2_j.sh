# Round 2, call J: fused MBConv front with the column-pair depthwise phase (tests + A/B).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -x -k "mbconv" 2>&1 | tail -6
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_bench_shapes.py -q -m gpu --timeout 600 -x 2>&1 | tail -5
for ff in 1 0; do
  EDET_FUSE_FRONT=$ff timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/j_ops_d0_ff$ff.json > gpurun_out/j_bench_d0_ff$ff.log 2>&1
  echo "fuse_front=$ff: $(tail -1 gpurun_out/j_bench_d0_ff$ff.log | cut -c1-230)"
done
python - <<'PY'
import json
d = json.load(open('gpurun_out/j_ops_d0_ff1.json'))
print([(r['name'], round(r['ms'] * 1e3, 1)) for r in d['ops'] if 'expand_dw' in r['name']])
d = json.load(open('gpurun_out/j_ops_d0_ff0.json'))
print([(r['name'], round(r['ms'] * 1e3, 1)) for r in d['ops'] if r['name'] in ('blocks_1/expand', 'blocks_1/dw', 'blocks_5/expand', 'blocks_5/dw')])
PY
