# Round 2, call M: pipeline variants (stream priorities, held-back head stage, grid slack), e2e
# with three requests in flight.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_serving.py -q -m gpu --timeout 600 -x 2>&1 | tail -8 | tee gpurun_out/m_tests.log
b() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ${CFG:+--config $CFG} > gpurun_out/m_bench_$name.log 2>&1
  python - "$name" <<'P'
import json,sys
name=sys.argv[1]
try:
  l=[x for x in open('gpurun_out/m_bench_%s.log'%name) if x.startswith('{')][-1]; d=json.loads(l)
  print('%-22s value %8.1f (%.3f ms)  e2e %8.1f (%.3f ms)'%(name,d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step']))
except Exception as e:
  print(name,'FAILED',e); print(open('gpurun_out/m_bench_%s.log'%name).read()[-600:])
P
}
b default A=1
b nmsprio EDET_NMS_PRIO=1
b defer EDET_DEFER_HEADS=1
b defer_nmsprio EDET_DEFER_HEADS=1 EDET_NMS_PRIO=1
b defer_headprio EDET_DEFER_HEADS=1 EDET_HEAD_PRIO=1 EDET_NMS_PRIO=1
b slack16 EDET_PERSIST_SLACK=16
b defer_slack16 EDET_DEFER_HEADS=1 EDET_PERSIST_SLACK=16
b nopipe EDET_PIPELINE=0
CFG=d4 b d4_default A=1
CFG=d4 b d4_defer EDET_DEFER_HEADS=1
CFG=d7x b d7x_default A=1
CFG=d7x b d7x_defer EDET_DEFER_HEADS=1
