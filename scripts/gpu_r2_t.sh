# Round 2, call T: dw_tile SE partials without shared-memory atomics; kernel + network tests, bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_network.py -q -m gpu --timeout 600 -x 2>&1 | tail -6 | tee gpurun_out/t_tests.log
for i in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/t_bench_$i.log 2>&1
python - $i <<'P'
import json,sys
l=[x for x in open('gpurun_out/t_bench_%s.log'%sys.argv[1]) if x.startswith('{')][-1]; d=json.loads(l); k=d['roofline']['per_kind']
print('value %8.1f (%.3f ms)  e2e %8.1f (%.3f ms)'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step']), {n:round(v['ms'],3) for n,v in k.items() if n.startswith('depthwise')})
P
done
