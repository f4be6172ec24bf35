cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -x -k "depthwise" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/ops_r1k.json > gpurun_out/bench16.log 2>&1; tail -1 gpurun_out/bench16.log | cut -c1-330
python - <<'PY'
import json
d=json.load(open('gpurun_out/ops_r1k.json'))
print({k:v['ms'] for k,v in d['kinds'].items()})
PY
