# Round 2, call N: class head fused with the class arg-max (edet_class_argmax), bias from shared
# memory in pointwise_tc.  Whole GPU suite, then A/B bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -12 | tee gpurun_out/n_tests.log
b() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ${CFG:+--config $CFG} > gpurun_out/n_bench_$name.log 2>&1
  python - "$name" <<'P'
import json,sys
name=sys.argv[1]
try:
  l=[x for x in open('gpurun_out/n_bench_%s.log'%name) if x.startswith('{')][-1]; d=json.loads(l)
  k=d['roofline']['per_kind']
  print('%-22s value %8.1f (%.3f ms)  e2e %8.1f (%.3f ms)  pw %.3f ms pre_nms %.3f ms'%(name,d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step'],k.get('pointwise_tc',{}).get('ms',0),k.get('pre_nms',{}).get('ms',0)))
except Exception as e:
  print(name,'FAILED',e); print(open('gpurun_out/n_bench_%s.log'%name).read()[-600:])
P
}
b default A=1
b nofuse EDET_FUSE_ARGMAX=0
b default2 A=1
b nopipe EDET_PIPELINE=0
CFG=d4 b d4_default A=1
CFG=d4 b d4_nofuse EDET_FUSE_ARGMAX=0
CFG=d7x b d7x_default A=1
CFG=v2s b v2s_default A=1
