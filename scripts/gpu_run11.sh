set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 120 -x -k "sepconv" 2>&1 | tail -15
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 -x 2>&1 | tail -5
timeout 600 python scripts/ab_engine.py base=fuse_mbconv_front:0,fuse_sepconv:0 towers=fuse_mbconv_front:0,fuse_sepconv:1 all=fuse_mbconv_front:0,fuse_sepconv:1,fuse_sepconv_nodes:1 2>&1 | tail -5
