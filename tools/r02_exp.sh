set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "sigma_store_paths" > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02j_pytest.log
L=distributed_sigmoid_loss_b200/libsiglip_b200.so
for shape in "16384 1024" "8192 768" "4096 768"; do
  set -- $shape
  timeout 300 python tools/ab_r1_vs_r2.py --r1-lib $L --r2-env SIGLIP_GSTORE_DIRECT=2 --B $1 --D $2 --rounds 6 --block-ms 150 > gpurun_out/r02j_ab_direct2_$1x$2.log 2>&1
  tail -3 gpurun_out/r02j_ab_direct2_$1x$2.log
done
