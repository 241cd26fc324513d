set -x
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/multi_gpu_check.py 2>&1 | grep -v "^W0\|^\*\*\*\|UserWarning\|Consider using\|loss=abs" > gpurun_out/final_multi_gpu_check_n$N.log
tail -3 gpurun_out/final_multi_gpu_check_n$N.log
grep -c " OK" gpurun_out/final_multi_gpu_check_n$N.log; grep -c "FAIL" gpurun_out/final_multi_gpu_check_n$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/final_bench_n$N.err | tee gpurun_out/final_bench_n$N.json | cut -c1-200
if [ "$N" = "8" ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 4 --steps 20 --warmup 5 2>gpurun_out/final_bench_n4.err | tee gpurun_out/final_bench_n4.json | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29614 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 2>/dev/null | tee gpurun_out/final_bench_ref_n8.json | cut -c1-200
fi
