"""Per-SASS-region summary of an ncu report's source page: python scripts/ncu_src.py report.ncu-rep [chunk] [warps]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 100
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if 'Address' in r)
hdr = rows[hi]
isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stall_cols = [j for j, h in enumerate(hdr) if h.startswith('stall_') ]
data = [(r[isrc].strip(), int(r[isamp]), int(r[iex]), r) for r in rows[hi + 1:] if len(r) == len(hdr)]
tot_s = sum(d[1] for d in data); tot_e = sum(d[2] for d in data)
nw = max(d[2] for d in data[:5]) or 1
print(len(data), 'sass instrs; samples', tot_s, 'executed', tot_e, 'warps', nw)
for c in range(0, len(data), chunk):
    seg = data[c:c + chunk]
    ops = {}
    for s, _, e, _ in seg:
        t = s.split()
        op = (t[1] if s.startswith('@') else t[0]).split('.')[0]
        ops[op] = ops.get(op, 0) + e
    top = sorted(ops.items(), key=lambda x: -x[1])[:6]
    st = {}
    for _, _, _, r in seg:
        for j in stall_cols:
            try: st[hdr[j]] = st.get(hdr[j], 0) + int(r[j])
            except ValueError: pass
    stt = sorted(st.items(), key=lambda x: -x[1])[:3]
    ss = sum(d[1] for d in seg)
    if ss == 0 and sum(d[2] for d in seg) == 0: continue
    print(f'{c:5d} samp {100*ss/tot_s:5.1f}% exec {100*sum(d[2] for d in seg)/tot_e:5.1f}% |', ' '.join(f'{k}:{v//nw}' for k, v in top), '|', ' '.join(f'{k[6:]}:{100*v//max(1,tot_s)}' for k, v in stt))
