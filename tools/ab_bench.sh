#!/bin/bash
# A/B bench of build variants on one GPU: tools/ab_bench.sh "<lib[:ENV=VAL,...]> ..." [bench args]
# prints one line per variant: name, patches/s (device-resident), ms per step
specs="$1"; shift
mkdir -p gpurun_out
for spec in $specs; do
  lib="${spec%%:*}"; envs=""
  if [[ "$spec" == *:* ]]; then envs="$(echo "${spec#*:}" | tr ',' ' ')"; fi
  out=gpurun_out/ab_$(echo "$spec" | tr ':=,/' '____').json
  env PMVS_LIB=$PWD/pais-mvs_b200/lib/$lib.so $envs python bench.py --no-cpu-baseline --no-e2e "$@" > $out 2> $out.err
  python - "$spec" $out <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[2]).read().strip().split("\n")[-1])
    print("%-44s %10.1f patches/s  %8.2f ms/step  frac %.3f  kept %.3f evals %.1f" % (sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["frac"], d["config"]["converged_kept_fraction"], d["config"]["evaluations_per_patch"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
    print(open(sys.argv[2]+".err").read()[-2000:])
PY
done
