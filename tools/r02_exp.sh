set -x
L=distributed_sigmoid_loss_b200/libsiglip_b200.so
for shape in "16384 1024" "8192 768"; do
  set -- $shape
  timeout 300 python tools/ab_r1_vs_r2.py --r1-lib $L --r2-opts 6=2 --B $1 --D $2 --rounds 6 --block-ms 200 > gpurun_out/r02l_ab_mcast2_$1x$2.log 2>&1
  tail -3 gpurun_out/r02l_ab_mcast2_$1x$2.log
done
