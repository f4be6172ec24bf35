# Round 2, engine-level A/B switches re-measured on the final kernels (no code change).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
b() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ${CFG:+--config $CFG} > gpurun_out/ab_bench_$name.log 2>&1
  python - "$name" <<'P'
import json,sys
name=sys.argv[1]
try:
  l=[x for x in open('gpurun_out/ab_bench_%s.log'%name) if x.startswith('{')][-1]; d=json.loads(l)
  print('%-22s value %8.1f (%.3f ms)  e2e %8.1f (%.3f ms)'%(name,d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step']))
except Exception as e:
  print(name,'FAILED',e); print(open('gpurun_out/ab_bench_%s.log'%name).read()[-600:])
P
}
b default A=1
b defer EDET_DEFER_HEADS=1
b nmsprio EDET_NMS_PRIO=1
b defer_nmsprio EDET_DEFER_HEADS=1 EDET_NMS_PRIO=1
b nopipe EDET_PIPELINE=0
b smem113 EDET_PW_SMEM_KB=113
b default2 A=1
