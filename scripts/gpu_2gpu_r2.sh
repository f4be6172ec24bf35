# Round 2, 2 GPUs: the N > 1 path of bench.py exactly as the driver launches it.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/k_bench_2gpu.log 2>&1
grep '^{' gpurun_out/k_bench_2gpu.log | tail -1 | cut -c1-900
tail -3 gpurun_out/k_bench_2gpu.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/k_bench_2gpu_ref.log 2>&1
grep '^{' gpurun_out/k_bench_2gpu_ref.log | tail -1 | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --config d4 --steps 10 --warmup 3 > gpurun_out/k_bench_2gpu_d4.log 2>&1
grep '^{' gpurun_out/k_bench_2gpu_d4.log | tail -1 | cut -c1-400
