#!/bin/bash
# One GPU-box pass that produces the round's test log and bench lines: GPU tests, bench lines (both arms). The ncu captures
# (launch list, --set full of the default launch -> tools/ncu_profile_json.py -> profiles/r2_profile.json) are separate calls:
#   ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/<tag>_launches.csv \
#       python bench.py --patches 16384 --steps 2 --warmup 1 --no-cpu-baseline
#   ncu --set full --import-source on --clock-control none -k regex:refine_kernel -s 1 -c 1 -f -o gpurun_out/prof_<tag>_final \
#       python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e
# usage (on the GPU box, from the repo root): tools/final_measure.sh <tag>      e.g. r2
tag=${1:-r2}
mkdir -p gpurun_out
(python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/${tag}_pytest_gpu.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_reference_arm.json 2> gpurun_out/bench_${tag}_reference_arm.err
python bench.py > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err
cat gpurun_out/${tag}_pytest_gpu.log
cut -c1-700 gpurun_out/bench_${tag}_n1.json
cut -c1-300 gpurun_out/bench_${tag}_reference_arm.json
