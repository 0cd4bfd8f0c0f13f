#!/bin/bash
# Per-launch device time of refine_kernel inside a `tmvs -r` run (5 x 1600x1200), next to the batch size of each call
# (PMVS_DEBUG trace): where the driver's GPU seconds go. usage (GPU box, repo root): tools/tmvs_launches.sh
mkdir -p gpurun_out
D=/tmp/tmvs_scale_ds
TMVS_SCALE_DIR=$D python tools/tmvs_scale.py 1600 1200 1024 4 > /dev/null 2>&1
cd $D
PMVS_DEBUG=1 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:refine_kernel --csv --log-file $OLDPWD/gpurun_out/tmvs_launches.csv \
    $OLDPWD/pais-mvs_b200/bin/tmvs -r scene.nvm --config config.txt --out-dir $D > $OLDPWD/gpurun_out/tmvs_launches.out 2> $OLDPWD/gpurun_out/tmvs_launches.err
cd $OLDPWD
python - <<'PY'
import csv, re
n = [int(m.group(1)) for m in re.finditer(r"refine_launch: n (\d+) NW", open("gpurun_out/tmvs_launches.err").read())]
rows = [r for r in csv.reader(open("gpurun_out/tmvs_launches.csv")) if len(r) > 5 and r[0].isdigit()]
t = [float(r[-1].replace(",", "")) for r in rows]
unit = rows[0][-2] if rows else "?"
scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "msecond": 1.0, "ms": 1.0}.get(unit, 1e-6)
t = [v * scale for v in t]
print("launches %d (trace %d), kernel time %.3f s, patches %d" % (len(t), len(n), sum(t) / 1e3, sum(n)))
if len(n) == len(t):
    for lo, hi in ((0, 150), (150, 300), (300, 600), (600, 900), (900, 1200), (1200, 1e9)):
        sel = [(a, b) for a, b in zip(n, t) if lo <= a < hi]
        if sel:
            print("  n in [%4d, %5s): %3d calls, %7d patches, %8.1f ms, %7.1f patches/ms" % (lo, "inf" if hi > 1e8 else int(hi), len(sel), sum(a for a, _ in sel), sum(b for _, b in sel), sum(a for a, _ in sel) / sum(b for _, b in sel)))
PY
grep -E "gpu_seconds|phase" gpurun_out/tmvs_launches.out
