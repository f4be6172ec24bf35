# Round 2, call C: the TMA-tiled depthwise kernel and the three-team pointwise epilogue: kernel
# tests, A/B measurements, network tests, bench, segment timing.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_tf_semantics.py -q -m gpu --timeout 300 -x -k "depthwise or se_fc or pointwise" 2>&1 | tail -15 | tee gpurun_out/c_tests_kernels.log
timeout 300 python scripts/ab_depthwise.py gpurun_out/c_ab_depthwise.json 2>&1 | tail -20
timeout 900 python -m pytest tests/test_gpu_network.py tests/test_gpu_bench_shapes.py tests/test_effnetv2.py -q -m gpu --timeout 600 -x 2>&1 | tail -8 | tee gpurun_out/c_tests_net.log
for cfg in "0 0" "1 0" "0 2" "0 3" "1 2"; do
  set -- $cfg
  EDET_DW_IMPL=$1 EDET_PW_TEAMS=$2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/c_ops_d0_dw$1_pw$2.json > gpurun_out/c_bench_d0_dw$1_pw$2.log 2>&1
  echo "dw_impl=$1 pw_teams=$2: $(tail -1 gpurun_out/c_bench_d0_dw$1_pw$2.log | cut -c1-260)"
done
timeout 300 python scripts/time_segments.py d0 gpurun_out/c_segments_d0.json 2>&1 | tail -4
