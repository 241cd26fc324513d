set -x
export SIGLIP_PEER_TIMEOUT_MS=30000
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29611 tools/multi_gpu_check.py 2>&1 | grep -v "^W0\|^\*\*\*\|UserWarning\|Consider using\|loss=abs\|OMP_NUM" > gpurun_out/r02f_multi_gpu_check_n$N.log
tail -2 gpurun_out/r02f_multi_gpu_check_n$N.log; grep -c " OK" gpurun_out/r02f_multi_gpu_check_n$N.log; grep -c "FAIL" gpurun_out/r02f_multi_gpu_check_n$N.log
timeout 300 $TR --master-port 29613 bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/r02f_bench_n$N.err > gpurun_out/r02f_bench_n$N.json
python - <<PY
import json
for f in ("gpurun_out/r02f_bench_n$N.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "NO LINE", e); continue
    r=d["roofline"]
    print(f.split("/")[-1], "ms/step %.4f"%d["ms_per_step"], "burst %.4f"%d["burst"]["ms_per_step"], "tflops %.0f"%d["tflops_per_gpu"], "grad %.4f loss %.4f"%(r["avg_launch_ms"], r["loss_kernel"]["avg_launch_ms"]), "launches", d["gpu_launches"], "ws GiB %.2f"%(d["workspace_bytes"]/2**30), "parity", d["parity"] and d["parity"]["pass"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["with_grads"]["ms_per_step"])
    print("   per_rank", d["per_rank"]["ms_per_step_by_rank"], d["per_rank"]["kernel_ms_per_step_by_rank"], d["clocks"].get("per_rank_sm_mhz"))
    if "scaling_diag" in d: print("   diag", {k:v for k,v in d["scaling_diag"].items() if k not in ("what","note")})
    if d.get("parity"): print("   parity", d["parity"]["fused_fp32"])
    print("   nvlink", d.get("nvlink"))
PY
tail -c 300 gpurun_out/r02f_bench_n$N.err
