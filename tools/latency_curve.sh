#!/bin/bash
# Small-batch behaviour of pmvs_refine_batch (the host driver's expansion rounds are a few hundred candidates per call):
# ms per call and patches/s against the batch size, and the warps-per-patch choices at one frontier-sized batch.
# usage (GPU box, repo root): tools/latency_curve.sh
mkdir -p gpurun_out
for n in 74 148 296 444 592 1184 2368 4736; do
  python bench.py --patches $n --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('n %5d  %8.3f ms/call  %9.1f patches/s  e2e %9.1f' % ($n, d['ms_per_step'], d['value'], d['e2e']['value']))"
done
for nw in 5 8 16; do
  PMVS_NW=$nw python bench.py --patches 296 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('n   296 NW $nw  %8.3f ms/call  %9.1f patches/s' % (d['ms_per_step'], d['value']))"
done
