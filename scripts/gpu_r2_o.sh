# Round 2, call O: fused-argmax / serve_stream tests, A/B of the fused MBConv front and the fused
# BiFPN node kernel under the pipelined step, config-3 e2e through serve_stream.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_effnetv2.py tests/test_gpu_serving.py -q -m gpu --timeout 600 -x 2>&1 | tail -8 | tee gpurun_out/o_tests.log
b() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline ${CFG:+--config $CFG} > gpurun_out/o_bench_$name.log 2>&1
  python - "$name" <<'P'
import json,sys
name=sys.argv[1]
try:
  l=[x for x in open('gpurun_out/o_bench_%s.log'%name) if x.startswith('{')][-1]; d=json.loads(l)
  print('%-22s value %8.1f (%.3f ms)  e2e %8.1f (%.3f ms)'%(name,d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step']))
except Exception as e:
  print(name,'FAILED',e); print(open('gpurun_out/o_bench_%s.log'%name).read()[-600:])
P
}
b default A=1
b fuse_front EDET_FUSE_FRONT=1
b fuse_nodes EDET_FUSE_NODES=1
b teams2 EDET_PW_TEAMS=2
b dw_reg EDET_DW_IMPL=1
CFG=v2s b v2s_default A=1
CFG=d7x b d7x_fuse_nodes EDET_FUSE_NODES=1
