# Serving-path check: serving + network tests, two bench lines (value / e2e).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_serving.py tests/test_gpu_network.py -q -m gpu --timeout 600 -x 2>&1 | tail -4
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/serving_bench_$i.log 2>&1
python - $i <<'P'
import json,sys
l=[x for x in open('gpurun_out/serving_bench_%s.log'%sys.argv[1]) if x.startswith('{')][-1]; d=json.loads(l)
print('value %8.1f (%.3f ms)  e2e %8.1f (%.3f ms)'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step']))
P
done
