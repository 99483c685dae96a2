#!/bin/bash
# A/B of the trailing-halo mechanisms at N GPUs (developer tool): block means (default) vs whole frame by peer memory.
# usage: tools/ab_halo.sh N "c4 c5" "means peer"
N=${1:-2}; port=29601
for m in ${3:-means peer}; do for w in ${2:-c4 c5}; do
  port=$((port + 1))
  PG_HALO=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $N --steps 10 --warmup 3 --workload $w --no-cpu --skip-e2e --skip-variants \
      > gpurun_out/ab_${w}_n${N}_${m}.json 2> gpurun_out/ab_${w}_n${N}_${m}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_${w}_n${N}_${m}.json"))
    print("$w n$N $m", round(d["ms_per_step"], 3), "ms/step", "%.4g" % d["value"], "k1_ms", round(d["roofline"]["k1_ms"], 3), d["config"]["parallelism"], d["clocks"]["sm_mhz"])
except Exception as e:
    print("$w n$N $m FAILED", e)
PY
done; done
