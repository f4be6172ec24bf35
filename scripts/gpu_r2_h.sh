# Round 2, call H: pointwise accumulator stages / team rotation, stem cp.async pipeline.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -x -k "stem or pointwise" 2>&1 | tail -6
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_bench_shapes.py tests/test_effnetv2.py -q -m gpu --timeout 600 -x 2>&1 | tail -5
for c in d0 d4 d7x v2s; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/h_ops_$c.json > gpurun_out/h_bench_$c.log 2>&1
  echo "$c: $(tail -1 gpurun_out/h_bench_$c.log | cut -c1-230)"
done
python - <<'PY'
import json
d = json.load(open('gpurun_out/h_ops_d0.json'))
print([(r['name'], round(r['ms'] * 1e3, 1)) for r in d['ops'] if r['name'] in ('stem', 'blocks_0/project', 'blocks_1/project', 'blocks_2/project', 'blocks_9/project')])
PY
