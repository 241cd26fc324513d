# ncu evidence for the final round-2 binary (one GPU): launch list of a bench run, --set full of the two kernels at the
# headline shape, then plain bench lines of the other single-GPU BASELINE shapes
set -x
NCU=/usr/local/cuda/bin/ncu
timeout 500 $NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches.csv \
   python bench.py --steps 4 --warmup 3 --sustain-ms 20 --no-cpu-baseline --no-parity > gpurun_out/r02f_launches_bench.log 2>&1
timeout 500 $NCU --set full --clock-control none --import-source on -k regex:siglip_gemm -s 4 -c 2 -f -o gpurun_out/r02f_full \
   python tools/profile_target.py --iters 3 > gpurun_out/r02f_full.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
timeout 300 python bench.py --batch 4096 --dim 768 --api fused --no-cpu-baseline > gpurun_out/r02f_bench_n1_b4096_d768_fused.json 2>/dev/null
timeout 300 python bench.py --batch 8192 --dim 768 --api fused --no-cpu-baseline > gpurun_out/r02f_bench_n1_b8192_d768_fused.json 2>/dev/null
timeout 300 python bench.py --batch 32768 --dim 1152 --no-cpu-baseline > gpurun_out/r02f_bench_n1_b32768_d1152_module.json 2>/dev/null
timeout 600 python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02f_bench_n1*.json")):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,"NO LINE",e); continue
    r=d["roofline"]; e=d["e2e"]
    print(f.split("/")[-1], "ms/step %.4f burst %.4f grad %.4f loss %.4f frac %.3f e2e %.4f parity %s clocks %s" % (
      d["ms_per_step"], d["burst"]["ms_per_step"], r["avg_launch_ms"], r["loss_kernel"]["avg_launch_ms"], r["frac"], e["ms_per_step"], d["parity"] and d["parity"]["pass"], d["clocks"]["sm_mhz"]))
PY
