"""Stall samples of an ncu source page, in segments of N SASS lines: executions, share of the samples, opcode mix, top stall
reasons.  usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_segments.py src.csv [lines per segment]"""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, ie, iss = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    try: n, s = int(r[ie]), int(r[iss])
    except: continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
    op = m.group(2).split(".")[0] if m else "?"
    st = {h: int(r[i]) for i, h in stall_cols if r[i].isdigit()}
    data.append((n, s, op, st, r[ia].strip()))
tot_s = sum(d[1] for d in data)
# segments: split when exec count changes by >25% and sustained
segs = []; start = 0
W = int(sys.argv[2]) if len(sys.argv) > 2 else 150
for a in range(0, len(data), W):
    seg = data[a:a+W]
    n = sum(d[0] for d in seg); s = sum(d[1] for d in seg)
    ops = collections.Counter()
    for d in seg: ops[d[2]] += d[0]
    st = collections.Counter()
    for d in seg:
        for k, v in d[3].items(): st[k] += v
    print("%5d-%5d exec/line %7.0fk samples %5.1f%% cyc/inst %5.2f | %s | %s" % (a, a+len(seg), n/len(seg)/1e3, 100*s/tot_s, (s/tot_s)/(n/sum(d[0] for d in data)) if n else 0,
          " ".join("%s %.0f%%" % (o, 100*c/n) for o, c in ops.most_common(5)) if n else "", " ".join("%s %.0f%%" % (k.replace("stall_",""), 100*v/max(1,s)) for k, v in st.most_common(4))))
