"""Aggregate an `ncu --page source --csv` dump: samples by opcode, by code region between barriers, and the hottest instructions."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) > 5]
si = hdr.index("Source"); ai = hdr.index("Warp Stall Sampling (All Samples)"); ei = hdr.index("Instructions Executed")
I = lambda x: int(x) if x.strip().isdigit() else 0
tot = sum(I(r[ai]) for r in data); totI = sum(I(r[ei]) for r in data)
def opof(s):
    t = s.split()
    if not t: return "?"
    op = t[1] if t[0].startswith('@') and len(t) > 1 else t[0]
    return op.split('.')[0]
byop = collections.Counter(); byopI = collections.Counter()
for r in data:
    byop[opof(r[si])] += I(r[ai]); byopI[opof(r[si])] += I(r[ei])
print("SASS instructions", len(data), " samples", tot, " warp-instr executed", totI)
for op, c in byop.most_common(16): print("  %-10s samples %7d (%4.1f%%)   executed %11d (%4.1f%%)" % (op, c, 100*c/max(tot,1), byopI[op], 100*byopI[op]/max(totI,1)))
bars = [i for i, r in enumerate(data) if 'BAR' in r[si]]
print("barriers at SASS index", bars)
edges = [0] + bars + [len(data)]
for a, b in zip(edges[:-1], edges[1:]):
    s = sum(I(r[ai]) for r in data[a:b]); e = sum(I(r[ei]) for r in data[a:b])
    print("  region [%5d,%5d)  samples %7d (%4.1f%%)  executed %11d (%4.1f%%)" % (a, b, s, 100*s/max(tot,1), e, 100*e/max(totI,1)))
print("hottest instructions:")
for i, r in sorted(enumerate(data), key=lambda ir: -I(ir[1][ai]))[:14]:
    print("  [%5d] %6d  %s" % (i, I(r[ai]), r[si].strip()[:90]))
