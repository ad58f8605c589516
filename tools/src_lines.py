"""Warp-stall samples of an ncu report aggregated by CUDA source line: python tools/src_lines.py rep.ncu-rep [kernel-regex] [n]"""
import csv, subprocess, sys
rep = sys.argv[1]
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"]
if len(sys.argv) > 2 and sys.argv[2]:
    cmd += ["--kernel-name", "regex:" + sys.argv[2]]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rows = list(csv.reader(subprocess.run(cmd, capture_output=True, text=True).stdout.splitlines()))
out, cur = [], "?"
for r in rows:
    if r and r[0] == "File Path" and len(r) > 1:
        cur = r[1].split("/")[-1]
    if len(r) > 5 and r[0].isdigit():
        try:
            s = int(r[4] or 0)
        except ValueError:
            s = 0
        out.append((s, cur, int(r[0]), r[1].strip()[:120]))
T = max(sum(o[0] for o in out), 1)
print("total samples", T)
for s, f, l, src in sorted(out, reverse=True)[:n]:
    print(f"{100*s/T:5.1f}% {f}:{l}: {src}")
