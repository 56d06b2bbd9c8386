"""Rewrites the 'Last bench line of the round' table of profiles/README.md from profiles/bench_r01_latest.json."""
import json, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
l = json.load(open(os.path.join(ROOT, "profiles", "bench_r01_latest.json")))
fp = l["frame_pipeline"]
rows = [
    ("`value` (device-resident, C2)", f"{l['value']/1e6:.2f} M solves/s, {l['ms_per_step']:.3f} ms per 10 000 pairs"),
    ("`e2e` (pinned host buffers through `pnec_solve_batch`)",
     f"{l['e2e']['value']/1e3:.0f} k solves/s, {l['e2e']['ms_per_step']:.2f} ms per step = "
     f"{l['e2e']['h2d_bytes_per_step']/l['e2e']['ms_per_step']/1e6:.1f} GB/s of H2D: PCIe-bound"),
    ("`roofline` K1", f"{l['roofline']['achieved']:.0f} GB/s = {l['roofline']['frac']:.3f} of measured peak ({l['roofline']['peak']} GB/s)"),
    (f"`cpu_baseline` (oracle, {l['cpu_baseline']['cores']} cores)",
     f"{l['cpu_baseline']['value']:.0f} solves/s ({l['cpu_baseline']['value_1core']:.0f} on one core)"),
    ("`frame_pipeline` (whole `PNEC::Solve`, no RANSAC)",
     f"{fp['value']/1e6:.2f} M frame pairs/s, {fp['ms_per_step']:.2f} ms per 10 000 pairs, {fp['gpu_launches_per_step']} launches; "
     f"oracle restatement {fp['cpu_baseline']['value']:.0f} pairs/s on {fp['cpu_baseline']['cores']} cores"),
    ("clocks during the timed region", f"{l['clocks']['sm_mhz']:.0f} MHz of {l['clocks']['sm_max_mhz']:.0f}, reasons {l['clocks']['reasons']}"),
]
table = "| quantity | value |\n|---|---|\n" + "\n".join(f"| {a} | {b} |" for a, b in rows) + "\n"
p = os.path.join(ROOT, "profiles", "README.md")
s = open(p).read()
s2 = re.sub(r"(## Last bench line of the round \(`bench_r01_latest.json`\)\n\n)\| quantity \| value \|\n\|---\|---\|\n(?:\|.*\n)+", lambda m: m.group(1) + table, s)
assert s2 != s or table in s
open(p, "w").write(s2)
print(table)
