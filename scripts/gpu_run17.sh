cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 300 python scripts/ab_engine.py base= 2>&1 | tail -3
EDET_PW_SPLIT=1 timeout 300 python scripts/ab_engine.py split= 2>&1 | tail -3
EDET_PW_SPLIT=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -x -k "pointwise" 2>&1 | tail -3
