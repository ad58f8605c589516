"""Per-kernel counts of the tensor-core / TMA / TMEM / packed-fp32 SASS opcodes in libisp_b200.so (cuobjdump -sass):
  python tools/sass_opcodes.py > profiles/rNN_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "isegprobe_b200", "libisp_b200.so")
OPS = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2",
       "MUFU.EX2", "REDG", "STL", "LDL"]
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        counts[fn] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn:
        op = m.group(1)
        counts[fn]["_total"] += 1
        for o in OPS:
            if op == o or op.startswith(o + "."):
                counts[fn][o] += 1
dem = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("# SASS opcode counts per kernel of isegprobe_b200/libisp_b200.so (sm_100a); kernels without any listed opcode omitted")
print("# columns: " + " ".join(OPS) + " | total instructions")
tot = collections.Counter()
for (fn, c), name in zip(counts.items(), dem):
    if not any(c[o] for o in OPS[:9] + OPS[9:12]):
        continue
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("isp::", "")
    print(f"{short[:90]:90s} " + " ".join(f"{c[o]:5d}" for o in OPS) + f" | {c['_total']}")
    tot.update(c)
print(f"{'TOTAL':90s} " + " ".join(f"{tot[o]:5d}" for o in OPS) + f" | {tot['_total']}")
