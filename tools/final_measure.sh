#!/bin/bash
# One GPU-box pass that produces everything profiles/ holds for a round: GPU tests, bench lines (both arms), the ncu
# launch list, one --set full capture of refine_kernel and its dram traffic at the default launch size.
# usage (on the GPU box, from the repo root): tools/final_measure.sh <tag>      e.g. r1
tag=${1:-r1}
mkdir -p gpurun_out
(python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/${tag}_pytest_gpu.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${tag}_reference_arm.json 2> gpurun_out/bench_${tag}_reference_arm.err
python bench.py > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --patches 2048 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:refine_kernel -s 1 -c 1 -f -o gpurun_out/prof_${tag}_final \
    python bench.py --patches 1184 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/prof_${tag}_final.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:refine_kernel -s 3 -c 1 --csv \
    --log-file gpurun_out/${tag}_traffic_refine_kernel.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_traffic.log 2>&1
cat gpurun_out/${tag}_pytest_gpu.log
cut -c1-400 gpurun_out/bench_${tag}_n1.json
cut -c1-300 gpurun_out/bench_${tag}_reference_arm.json
