# A/B of the current library against tools/_base_libsiglip_b200.so (built from the previous commit) + GPU tests
set -x
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02g_pytest_gpu.log
for shape in "4096 768" "8192 768" "16384 1024"; do
  set -- $shape
  timeout 300 python tools/ab_r1_vs_r2.py --r1-lib tools/_base_libsiglip_b200.so --B $1 --D $2 --rounds 6 --block-ms 150 > gpurun_out/r02g_ab_$1x$2.log 2>&1
  tail -4 gpurun_out/r02g_ab_$1x$2.log
done
timeout 200 python tools/launch_timeline.py --B 4096 --D 768 > gpurun_out/r02g_timeline_4096x768.log 2>&1
tail -n 3 gpurun_out/r02g_timeline_4096x768.log
