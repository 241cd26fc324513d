set -x
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 2>gpurun_out/final_bench.err | tee gpurun_out/final_bench_n1.json | cut -c1-300
python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tee gpurun_out/final_bench_ref.json | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 90 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 3 --warmup 3 --sustain-ms 0 --no-cpu-baseline > gpurun_out/final_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:siglip_gemm_kernel -s 4 -c 2 -o gpurun_out/final_full python tools/profile_target.py --iters 4 > gpurun_out/final_ncu_full.log 2>&1
ls -la gpurun_out/final_*
