"""Opcode histogram of the largest loop of a device function inside libpmvs_b200.so (no GPU needed).
usage: python tools/sass_loop.py <substring of the function label> [kernel substring [loop index]]   (SASS_DUMP=1 prints the loop)"""
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
so = os.path.join(ROOT, "pais-mvs_b200", "lib", "libpmvs_b200.so")
want = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else "refine_kernel"
with tempfile.TemporaryDirectory() as d:
    subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=d, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
lines = txt.split("\n")
start = [i for i, l in enumerate(lines) if l.endswith(":") and want in l and kern in l][0]
end = [i for i, l in enumerate(lines) if i > start and l.strip().startswith(".type")]
body = lines[start:end[0] if end else len(lines)]
labels = {}
for i, l in enumerate(body):
    m = re.match(r"^(\.L_x_\d+):", l)
    if m:
        labels[m.group(1)] = i
loops = []
for i, l in enumerate(body):
    m = re.search(r"BRA.*`\((\.L_x_\d+)\)", l)
    if m and m.group(1) in labels and labels[m.group(1)] < i:
        loops.append((labels[m.group(1)], i))
print("function lines", len(body), "loops", loops)
a, b = loops[int(sys.argv[3])] if len(sys.argv) > 3 else max(loops, key=lambda t: t[1] - t[0])
if os.environ.get("SASS_DUMP"):
    print("\n".join(body[a:b + 1]))
ops = collections.Counter()
for l in body[a:b + 1]:
    m = re.search(r"\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        ops[m.group(2)] += 1
tot = sum(ops.values())
f64 = sum(v for k, v in ops.items() if k.split(".")[0] in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"))
print("largest loop: %d instructions, %d fp64-pipe" % (tot, f64))
for k, v in ops.most_common(45):
    print("  %-24s %d" % (k, v))
