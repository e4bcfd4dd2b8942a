#!/bin/bash
# Multi-GPU bench lines: bash scripts/gpu_multi.sh <N> "<workload:scaling> ..."   (one JSON line each under gpurun_out/)
N=${1:-2}; shift
mkdir -p gpurun_out
for spec in ${@:-cfg2:weak cfg3:strong}; do
  W=${spec%%:*}; S=${spec##*:}
  P=$((29500 + RANDOM % 1000))
  out=gpurun_out/multi_${W}_${S}_${N}gpu.json
  if [ $N -gt 1 ]; then
    timeout ${BENCH_TIMEOUT:-400} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload $W --scaling $S --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline > $out 2> gpurun_out/multi_${W}_${S}_${N}gpu.err
  else
    timeout ${BENCH_TIMEOUT:-400} python bench.py --gpus 1 --workload $W --scaling $S --steps ${STEPS:-5} --warmup 3 --no-cpu-baseline > $out 2> gpurun_out/multi_${W}_${S}_${N}gpu.err
  fi
  echo "exit $? $spec N=$N"
  python - $out <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1]); c = d["config"]; r = d["roofline"]
    print(f"  {d['value']/1e6:9.3f} M/s  e2e {d['e2e']['value']/1e6:9.3f} M/s  ms/step {d['ms_per_step']:.2f}  frac {r['frac']:.3f}  global_rows {c['global_rows']}  "
          f"steps {c['dopri5_steps']} identical_across_ranks {c['steps_identical_across_ranks']} shard_check {c['shard_check']} numa {c['numa']} clocks {d['clocks']['sm_mhz']}")
except Exception as e:
    print("  FAILED", e); print(open(sys.argv[1].replace(".json", ".err")).read()[-1500:])
PY
done
