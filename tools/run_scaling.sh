#!/bin/bash
# Measurement rows of BASELINE.json configs[1..4] at N ranks of one box (run under `gpurun --gpus N`):
#   tools/run_scaling.sh N [tag]      -> gpurun_out/<tag>_<workload>_<n>gpu.json  for n = 1 (same-box baseline) and n = N
# Launch exactly as the driver does: python bench.py for N = 1, torch.distributed.run for N > 1.
N=${1:-1}; TAG=${2:-r02}
OUT=gpurun_out; mkdir -p $OUT
run() {  # n workload steps extra...
  local n=$1 w=$2 k=$3; shift 3
  local f=$OUT/${TAG}_${w}_${k}steps_${n}gpu.json
  if [ "$n" = 1 ]; then python bench.py --gpus 1 --workload $w --steps $k --no-cpu-baseline "$@" > $f 2> ${f%.json}.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --workload $w --steps $k "$@" > $f 2> ${f%.json}.err; fi
  echo "== $w N=$n steps=$k rc=$?"; tail -c 400 $f | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read().strip().splitlines()[-1]); e = d.get('e2e') or {}
    print('   value %.5g %s  ms/step %.4f  e2e %s  frac %.3f  clocks %s %s' % (d['value'], d['unit'], d['ms_per_step'], e.get('value'), d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons']))
except Exception as ex:
    print('   (no JSON line)', ex)
"
}
NS="1 $N"; [ "$N" = 1 ] && NS="1"
for n in $NS; do
  run $n cpn1024 20 --warmup 5
  run $n cpn1024 2000 --warmup 20
  run $n gt1024x5 200 --warmup 5
  run $n twostage 200 --warmup 5
  run $n evalloop 2000 --warmup 20
done
run $N sweep1m 1 --warmup 3
