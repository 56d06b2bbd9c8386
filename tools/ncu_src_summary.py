"""Summarise `ncu --page source --csv` output: instruction mix by opcode and top stall sites.
usage: python tools/ncu_src_summary.py report.ncu-rep [topN]"""
import csv, subprocess, sys, collections, io, re
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# find header row
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
def f(r, name):
    try: return float(r[col[name]])
    except Exception: return 0.0
tot_inst = sum(f(r, "Instructions Executed") for r in body)
tot_samp = sum(f(r, "# Samples") for r in body)
print(f"kernel: {rows[0][1] if rows[0] else ''}")
print(f"SASS lines {len(body)}  warp-instructions {tot_inst:.3e}  samples {tot_samp:.0f}")
by_op = collections.Counter(); samp_op = collections.Counter()
for r in body:
    src = r[col["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2).split(".")[0] if m else src[:10]
    by_op[op] += f(r, "Instructions Executed"); samp_op[op] += f(r, "# Samples")
print("\nopcode mix (warp instr %, stall-sample %):")
for op, c in by_op.most_common(22):
    print(f"  {op:10s} {100*c/tot_inst:6.2f}%   {100*samp_op[op]/max(tot_samp,1):6.2f}%")
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot_st = {h: sum(f(r, h) for r in body) for h in stall_cols}
s = sum(tot_st.values())
print("\nstall reasons (all samples):")
for h, v in sorted(tot_st.items(), key=lambda kv: -kv[1])[:10]:
    print(f"  {h:24s} {100*v/max(s,1):6.2f}%")
print(f"\ntop {topn} SASS lines by samples:")
for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:topn]:
    top = max(stall_cols, key=lambda h: f(r, h))
    print(f"  {f(r,'# Samples'):7.0f}  {f(r,'Instructions Executed'):10.0f}  {top:18s} {r[col['Source']].strip()[:90]}")
