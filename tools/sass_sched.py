"""Static schedule of the largest loop of a device function inside libpmvs_b200.so: every instruction with the stall
count, yield flag, scoreboard set/wait fields of its control word (Volta+ 128-bit encoding: stall = bits 105..108,
yield 109, write barrier 110..112, read barrier 113..115, wait mask 116..121). No GPU needed.
usage: python tools/sass_sched.py <substring of the function label> [kernel substring] [--all]"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
so = os.path.abspath(os.environ.get("PMVS_LIB", os.path.join(ROOT, "pais-mvs_b200", "lib", "libpmvs_b200.so")))
args = [a for a in sys.argv[1:] if not a.startswith("--")]
want = args[0]
kern = args[1] if len(args) > 1 else "refine_kernelILi256ELi2"
with tempfile.TemporaryDirectory() as d:
    subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=d, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "-c", "-hex", os.path.join(d, cubin)], capture_output=True, text=True).stdout
lines = txt.split("\n")
start = [i for i, l in enumerate(lines) if l.endswith(":") and want in l and kern in l][0]
end = [i for i, l in enumerate(lines) if i > start and l.strip().startswith(".type")]
body = lines[start:end[0] if end else len(lines)]
# join the two hex words of each instruction
insts = []   # (label or None, text, ctrlword)
i = 0
pending_label = None
while i < len(body):
    l = body[i]
    m = re.match(r"^(\.L_x_\d+):", l)
    if m:
        pending_label = m.group(1)
        i += 1
        continue
    m = re.match(r"^\s+/\*([0-9a-f]+)\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/", l)
    if m:
        hi = None
        if i + 1 < len(body):
            m2 = re.match(r"^\s+/\* (0x[0-9a-f]+) \*/", body[i + 1])
            if m2:
                hi = int(m2.group(1), 16)
                i += 1
        insts.append((pending_label, m.group(1), m.group(2).strip(), hi))
        pending_label = None
    i += 1
labels = {lab: k for k, (lab, _, _, _) in enumerate(insts) if lab}
loops = []
for k, (_, _, t, _) in enumerate(insts):
    m = re.search(r"BRA.*`\((\.L_x_\d+)\)", t)
    if m and m.group(1) in labels and labels[m.group(1)] < k:
        loops.append((labels[m.group(1)], k))
if "--all" in sys.argv:
    a, b = 0, len(insts) - 1
else:
    inner = [lp for lp in loops if not any(o != lp and lp[0] <= o[0] and o[1] <= lp[1] for o in loops)]
    a, b = max(inner, key=lambda t: t[1] - t[0])
print("loops", loops, "-> showing", (a, b))
tot_stall = 0
n = 0
for lab, addr, t, hi in insts[a:b + 1]:
    ctrl = (hi >> 41) & 0x7fffff if hi is not None else 0
    stall = ctrl & 0xf
    yld = (ctrl >> 4) & 1
    wb = (ctrl >> 5) & 7
    rb = (ctrl >> 8) & 7
    wm = (ctrl >> 11) & 0x3f
    tot_stall += max(stall, 1)
    n += 1
    print("%-8s %s s=%2d %s w%s r%s wait=%02x  %s" % (lab or "", addr, stall, "Y" if yld else "-", wb if wb != 7 else "-",
                                                     rb if rb != 7 else "-", wm, t[:110]))
print("instructions", n, "sum of stall counts", tot_stall)
