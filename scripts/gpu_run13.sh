set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -x -k "conv2d_tc" 2>&1 | tail -15
timeout 600 python -m pytest tests/test_effnetv2.py -q -m gpu --timeout 300 -s 2>&1 | tail -15
