# The ncu captures behind profiles/r02_* (run under gpurun; every command first exits 0 without ncu).
#   bash tools/profile_r02.sh [bench|variants]
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu --skip-e2e --skip-variants --skip-c5"
if [ "${1:-bench}" = "bench" ]; then
  $B > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv $B > gpurun_out/ncu_launches.log 2>&1
  $B > gpurun_out/plain_bench2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_tiled_b88 -s 3 -c 2 -f -o gpurun_out/prof_k1_r02 $B > gpurun_out/ncu_k1.log 2>&1
else
  for c in rich_b388_tiled_2folds:prof_rich_r02 adv_b388_tiled:prof_adv_r02; do
    Q="python tools/quick_bench.py --frames 256 --only ${c%%:*}"
    $Q > gpurun_out/plain_q.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_tiled_b88 -s 2 -c 1 -f -o gpurun_out/${c##*:} $Q > gpurun_out/ncu_q.log 2>&1
  done
  for c in ks_true_pointwise_tiled_2folds:prof_pw_ks_r02 basic_pointwise_tiled_2folds:prof_pw_basic_r02; do
    Q="python tools/quick_bench.py --pointwise --frames 256 --only ${c%%:*}"
    $Q > gpurun_out/plain_q.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k1_tiled_pw -s 2 -c 1 -f -o gpurun_out/${c##*:} $Q > gpurun_out/ncu_q.log 2>&1
  done
fi
