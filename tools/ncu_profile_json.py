"""profiles/r2_profile.json from an ncu --set full capture of refine_kernel (read here, no GPU):
    python tools/ncu_profile_json.py gpurun_out/prof.ncu-rep <patches of the captured launch> [mix_bound_frac]
bench.py copies issue_active_frac / fp64_pipe_frac / xu_pipe_frac / traffic into its `roofline` object."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
rep, patches = sys.argv[1], int(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
m = {h: v for h, v in zip(rows[0], rows[2])}
units = {h: u for h, u in zip(rows[0], rows[1])}


def val(key):
    hit = [h for h in m if h == key or h.endswith("." + key)]
    return float(m[hit[0]].replace(",", "")), units[hit[0]]


def bytes_of(key):
    v, u = val(key)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


out = {
    "capture": os.path.basename(rep),
    "kernel": m[[h for h in m if h.endswith("Kernel Name")][0]],
    "kernel_ms": val("gpu__time_duration.sum")[0] * {"ms": 1, "us": 1e-3, "s": 1e3, "ns": 1e-6}[val("gpu__time_duration.sum")[1]],
    "issue_active_frac": val("smsp__issue_active.avg.pct_of_peak_sustained_active")[0] / 100,
    "fp64_pipe_frac": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed")[0] / 100,
    "xu_pipe_frac": val("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active")[0] / 100,
    "warps_per_sm": val("sm__warps_active.avg.per_cycle_active")[0],
    "registers": val("launch__registers_per_thread")[0],
    "spill_instructions": val("sass__inst_executed_register_spilling")[0],
    "instructions": val("smsp__inst_executed.sum")[0],
    "l1_global_load_hit_frac": val("l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum")[0] / val("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")[0],
    "traffic": {"patches": patches, "dram_bytes_read": bytes_of("dram__bytes_read.sum"), "dram_bytes_write": bytes_of("dram__bytes_write.sum")},
}
if len(sys.argv) > 3:
    out["mix_bound_frac"] = float(sys.argv[3])
prev = {}
p = os.path.join(ROOT, "profiles", "r2_profile.json")
if os.path.exists(p):
    prev = json.load(open(p))
    for k in ("mix_bound_frac",):
        if k in prev and k not in out:
            out[k] = prev[k]
json.dump(out, open(p, "w"), indent=1)
print(json.dumps(out, indent=1))
