"""Attribute the warp-stall samples / executed instructions of an ncu report to the device functions inside
refine_kernel, using nvdisasm of the library that was profiled (must still be the one on disk).
usage: python tools/ncu_by_function.py report.ncu-rep [function-substring-to-detail]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
so = os.path.abspath(os.environ.get("PMVS_LIB", os.path.join(ROOT, "pais-mvs_b200", "lib", "libpmvs_b200.so")))
kern = os.environ.get("PMVS_KERNEL", "ILi256ELi2")     # template arguments of the profiled refine_kernel instantiation
rep = sys.argv[1]
detail = sys.argv[2] if len(sys.argv) > 2 else None
with tempfile.TemporaryDirectory() as d:
    subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=d, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text._Z13refine_kernel" + kern)][0]
end = [i for i, l in enumerate(txt) if i > start and l.startswith(".text.")]
end = end[0] if end else len(txt)
func, seq = "refine_kernel (main)", []
for l in txt[start:end]:
    if l.startswith("$_Z13refine_kernel"):
        func = re.sub(r"^_Z\d+", "", l.split("$")[2])[:44]
    elif l.startswith("$__internal"):
        func = l.strip(":")[:44]
    elif re.search(r"/\*[0-9a-f]{4,}\*/", l):
        seq.append(func)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) >= len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
if len(seq) != len(data):
    print("WARNING: library on disk (%d instr) is not the profiled one (%d instr)" % (len(seq), len(data)))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for k, r in enumerate(data):
    f = seq[k] if k < len(seq) else "?"
    agg[f][0] += int(r[ix["# Samples"]] or 0)
    agg[f][1] += int(r[ix["Instructions Executed"]] or 0)
    for h in stall_cols:
        agg[f][2][h[6:]] += int(r[ix[h]] or 0)
ts = sum(v[0] for v in agg.values())
ti = sum(v[1] for v in agg.values())
print("| function | samples | instructions | top stalls |\n|---|---|---|---|")
for f, v in sorted(agg.items(), key=lambda t: -t[1][0]):
    if v[0] < 0.002 * ts:
        continue
    st = ", ".join("%s %.0f%%" % (k, 100 * n / max(v[0], 1)) for k, n in v[2].most_common(4))
    print("| %s | %.1f%% | %.1f%% | %s |" % (f, 100 * v[0] / ts, 100 * v[1] / ti, st))
if detail:
    ks = [k for k in range(len(data)) if k < len(seq) and detail in seq[k]]
    top = sorted(ks, key=lambda k: -int(data[k][ix["# Samples"]] or 0))[:int(os.environ.get("TOPN", "40"))]
    print("\ntop stalled instructions in", detail)
    for k in sorted(top):
        r = data[k]
        st = sorted(((h[6:], int(r[ix[h]] or 0)) for h in stall_cols if int(r[ix[h]] or 0) > 0), key=lambda t: -t[1])
        print(k, r[ix["Source"]].strip()[:58], "n=%s s=%s" % (r[ix["Instructions Executed"]], r[ix["# Samples"]]), st[:2])
