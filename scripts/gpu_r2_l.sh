# Round 2, call L: pipelined engine (backbone of step i+1 under heads of step i), fp32 depthwise
# taps, format-model parity bars, pointwise smem budget.  Tests first, then A/B bench lines.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -x 2>&1 | tail -15 | tee gpurun_out/l_tests.log
b() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ${CFG:+--config $CFG} > gpurun_out/l_bench_$name.log 2>&1
  python - "$name" <<'P'
import json,sys
name=sys.argv[1]
try:
  l=[x for x in open('gpurun_out/l_bench_%s.log'%name) if x.startswith('{')][-1]; d=json.loads(l)
  print('%-22s value %8.1f (%.3f ms)  e2e %8.1f (%.3f ms)'%(name,d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step']))
except Exception as e:
  print(name,'FAILED',e); print(open('gpurun_out/l_bench_%s.log'%name).read()[-600:])
P
}
b default A=1
b nopipe EDET_PIPELINE=0
b noprio EDET_HEAD_PRIO=0
b slack16 EDET_PERSIST_SLACK=16
b slack32 EDET_PERSIST_SLACK=32
b smem113 EDET_PW_SMEM_KB=113
b nopipe_smem113 EDET_PIPELINE=0 EDET_PW_SMEM_KB=113
CFG=d4 b d4_default A=1
CFG=d4 b d4_nopipe EDET_PIPELINE=0
CFG=d7x b d7x_default A=1
CFG=d7x b d7x_nopipe EDET_PIPELINE=0
