N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for v in 1 0 1 0; do
SIGLIP_PDL=$v timeout 200 $TR --master-port 2962$v bench.py --gpus $N --steps 20 --warmup 5 --no-parity --no-scaling-diag 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('pdl=$v', 'A ms/step %.4f'%d['ms_per_step'], 'B(events) %.4f'%r['ms_per_step_with_kernel_events'], 'burst %.4f'%d['burst']['ms_per_step'], 'kernels', [round(x,3) for x in d['per_rank']['kernel_ms_per_step_by_rank']], 'clk', d['clocks']['per_rank_sm_mhz'], 'e2e %.3f'%d['e2e']['ms_per_step'])"
done
