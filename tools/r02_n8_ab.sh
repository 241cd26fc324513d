set -x
export SIGLIP_PEER_TIMEOUT_MS=30000
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29612 tools/scaling_diag.py --phases BCT --out gpurun_out/r02_scaling_diag_n$N.json 2>&1 | grep -v "^W0\|^\*\*\*\|OMP_NUM" > gpurun_out/r02_scaling_diag_n$N.log
grep -v "SM clock" gpurun_out/r02_scaling_diag_n$N.log | tail -40
timeout 200 $TR --master-port 29613 bench.py --gpus $N --steps 20 --warmup 5 --no-parity --no-scaling-diag 2>/dev/null > gpurun_out/r02_ab_n${N}_fused.json
timeout 200 $TR --master-port 29614 bench.py --gpus $N --steps 20 --warmup 5 --no-parity --no-scaling-diag --schedule split 2>/dev/null > gpurun_out/r02_ab_n${N}_split.json
SIGLIP_INKERNEL_SYNC=0 timeout 200 $TR --master-port 29615 bench.py --gpus $N --steps 20 --warmup 5 --no-parity --no-scaling-diag 2>/dev/null > gpurun_out/r02_ab_n${N}_fused_helperflags.json
timeout 200 $TR --master-port 29616 bench.py --gpus $N --steps 20 --warmup 5 --no-parity --no-scaling-diag 2>/dev/null > gpurun_out/r02_ab_n${N}_fused_2.json
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_ab_n${N}_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "NO LINE", e); continue
    r=d["roofline"]
    print(f.split("/")[-1], "ms/step %.4f"%d["ms_per_step"], "burst %.4f"%d["burst"]["ms_per_step"], "grad %.4f loss %.4f"%(r["avg_launch_ms"], r["loss_kernel"]["avg_launch_ms"]), "launches", d["gpu_launches"], "ws GiB %.2f"%(d["workspace_bytes"]/2**30), "clk", d["clocks"].get("per_rank_sm_mhz"))
PY
