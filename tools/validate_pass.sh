#!/bin/bash
# Short GPU-box pass after host-side changes: GPU tests, the default bench line, and `tmvs -r` / `tmvs -f` end to end at
# the named image size (5 x 1600x1200). usage (on the GPU box, from the repo root): tools/validate_pass.sh <tag>
tag=${1:-r1b}
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/${tag}_pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err
(timeout 400 python tools/tmvs_scale.py 1600 1200 1024 4 2>&1 | tail -20) > gpurun_out/${tag}_tmvs_1600x1200.log
cat gpurun_out/${tag}_pytest_gpu.log
cut -c1-600 gpurun_out/bench_${tag}_n1.json
cat gpurun_out/${tag}_tmvs_1600x1200.log
