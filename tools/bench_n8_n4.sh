for N in 8 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2962$N bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/final_bench_n$N.err | tee gpurun_out/final_bench_n$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], 'burst', d['burst']['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['sync_ms_per_step'], 'grad', d['roofline']['avg_launch_ms'], 'loss', d['roofline']['loss_kernel']['avg_launch_ms'], d['nvlink']['measured_rx_GBps'])"
done
