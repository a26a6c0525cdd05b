"""SASS opcode histogram of the shipped library, per kernel for the opcodes that prove which units run
(tensor cores: HMMA = mma.sync, IMMA = integer mma.sync, UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st; TMA: UBLKCP = cp.async.bulk;
packed fp32: FFMA2; fp64: DFMA):  python scripts/sass_hist.py > profiles/sass_r2.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "orthogonal_classifiers_b200/liboc_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
tot = collections.Counter()
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        depth = 0  # drop the parameter list (the last balanced parenthesis group), keep template arguments
        for i in range(len(cur) - 1, -1, -1):
            depth += cur[i] == ")"
            depth -= cur[i] == "("
            if depth == 0 and cur[i] == "(":
                cur = cur[:i]
                break
        cur = cur.replace("(int)", "").replace("(bool)", "").replace("void ", "")
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        op = m.group(1)
        tot[op] += 1
        per[cur][op] += 1
keys = ["FFMA", "FFMA2", "DFMA", "HMMA", "IMMA", "UTCHMMA", "LDTM", "STTM", "UBLKCP", "SHFL", "LDS", "STS", "LDG", "STG", "MUFU"]
print("SASS of", lib, "(sm_100a): instructions per kernel, selected opcodes")
print(f"{'kernel':58s}" + "".join(f"{k:>8s}" for k in keys) + f"{'total':>9s}")
for name, c in per.items():
    print(f"{name[-58:]:58s}" + "".join(f"{c.get(k, 0):8d}" for k in keys) + f"{sum(c.values()):9d}")
print(f"{'ALL':58s}" + "".join(f"{tot.get(k, 0):8d}" for k in keys) + f"{sum(tot.values()):9d}")
print()
print("whole-library histogram (top 40):")
for op, n in tot.most_common(40):
    print(f"  {op:12s} {n}")
