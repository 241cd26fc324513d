# Final single-GPU validation of the round-2 tree: GPU tests, smoke, default bench line, launch anatomy at two shapes
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02f_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; echo "smoke rc=$?"
tail -2 gpurun_out/r02f_smoke.log
timeout 600 python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02f_bench_n1.json").read().strip().splitlines()[-1])
r=d["roofline"]; e=d["e2e"]
print("ms/step %.4f burst %.4f grad %.4f loss %.4f frac %.3f e2e %.4f (first %.4f) grads %.4f sync %.4f parity %s clocks %s cpu %s" % (
 d["ms_per_step"], d["burst"]["ms_per_step"], r["avg_launch_ms"], r["loss_kernel"]["avg_launch_ms"], r["frac"], e["ms_per_step"],
 e["first_steps_ms_per_step"], e["with_grads"]["ms_per_step"], e["sync_ms_per_step"], d["parity"]["pass"], d["clocks"]["sm_mhz"], d["cpu_baseline"]["value"]))
PY
timeout 200 python tools/launch_timeline.py --B 16384 --D 1024 > gpurun_out/r02f_timeline_headline.log 2>&1
timeout 200 python tools/launch_timeline.py --B 4096 --D 768 > gpurun_out/r02f_timeline_4096x768.log 2>&1
tail -4 gpurun_out/r02f_timeline_headline.log gpurun_out/r02f_timeline_4096x768.log
