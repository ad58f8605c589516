"""Top stall locations of one kernel from an ncu report: python tools/src_stalls.py rep.ncu-rep [kernel-regex]"""
import csv, collections, subprocess, sys
rep = sys.argv[1]
cmd = ["ncu", "-i", rep, "--page", "source", "--csv"]
if len(sys.argv) > 2:
    cmd += ["--kernel-name", "regex:" + sys.argv[2]]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
i_src, i_s, i_ex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot, data = collections.Counter(), []
for n, r in enumerate(rows[2:]):
    if len(r) <= i_s or not r[i_s].strip().isdigit():
        continue
    s = int(r[i_s]); data.append((s, r[i_src].strip(), int(r[i_ex] or 0), n))
    for i, h in stall_cols:
        try: tot[h] += int(r[i] or 0)
        except ValueError: pass
T = max(sum(d[0] for d in data), 1)
print(rows[0][1][:100]); print("total samples", T, "instructions", len(data))
for h, c in tot.most_common(8): print(f"  {h:28s} {c:8d} {100*c/T:5.1f}%")
for s, src, ex, n in sorted(data, reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 24]:
    print(f"{100*s/T:5.1f}% {ex:10d} #{n:4d} {src[:100]}")
