# ncu evidence for round 2 (one GPU): launch list of a bench run, --set full of the two kernels (W=1 headline) and of
# a 3-chunk loopback step (gradient kernel with the fold + dimg accumulation active)
set -x
NCU=/usr/local/cuda/bin/ncu
timeout 500 $NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
   python bench.py --steps 4 --warmup 3 --sustain-ms 20 --no-cpu-baseline --no-parity > gpurun_out/r02_launches_bench.log 2>&1
timeout 500 $NCU --set full --clock-control none --import-source on -k regex:siglip_gemm -s 4 -c 2 -f -o gpurun_out/r02_full \
   python tools/profile_target.py --iters 3 > gpurun_out/r02_full.log 2>&1
timeout 500 $NCU --set full --clock-control none --import-source on -k regex:siglip_gemm -s 6 -c 6 -f -o gpurun_out/r02_full_w3 \
   python tools/profile_target.py --iters 2 --loopback 3 > gpurun_out/r02_full_w3.log 2>&1
ls -la gpurun_out/*.ncu-rep
