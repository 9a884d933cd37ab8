set -x
B="python bench.py --steps 64 --warmup 3 --skip-cpu --skip-other --e2e-steps 3"
$B > gpurun_out/r02_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r02_ncu_launches_bench.csv $B > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_n1_tma -s 2060 -c 2 -o gpurun_out/r02_step_n1_tma_full $B > gpurun_out/ncu_b.log 2>&1
for n in 8 64 256; do
  T="python tools/tiled_sweep.py --points $n:65536 --kin -1 --steps 32"
  $T > gpurun_out/plain_t$n.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_tiled -s 80 -c 1 -o gpurun_out/r02_tiled_n${n}_full $T > gpurun_out/ncu_t$n.log 2>&1
done
T="python tools/tiled_sweep.py --points 64:262144 --kin -1 --steps 32"
$T > gpurun_out/plain_t64b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_tiled -s 80 -c 1 -o gpurun_out/r02_tiled_n64_b262144_full $T > gpurun_out/ncu_t64b.log 2>&1
tail -2 gpurun_out/ncu_*.log
