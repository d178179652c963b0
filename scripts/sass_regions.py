"""Per-region (between barriers) stall breakdown and shared-memory wavefronts of an `ncu --page source --csv` dump."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) > 5]
half = len(data) // 2 if len(sys.argv) > 2 and sys.argv[2] == "half" else len(data)
data = data[:half]
si = hdr.index("Source"); ai = hdr.index("Warp Stall Sampling (All Samples)"); ei = hdr.index("Instructions Executed")
wi = hdr.index("L1 Wavefronts Shared"); wx = hdr.index("L1 Wavefronts Shared Excessive")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
I = lambda x: int(x) if x.strip().isdigit() else 0
tot = sum(I(r[ai]) for r in data)
bars = [i for i, r in enumerate(data) if 'BAR' in r[si]]
edges = [0] + bars + [len(data)]
for a, b in zip(edges[:-1], edges[1:]):
    seg = data[a:b]
    s = sum(I(r[ai]) for r in seg); e = sum(I(r[ei]) for r in seg)
    w = sum(I(r[wi]) for r in seg); x = sum(I(r[wx]) for r in seg)
    dp = sum(I(r[ei]) for r in seg if r[si].split()[0].startswith(('DMUL', 'DADD', 'DFMA')) or (len(r[si].split()) > 1 and r[si].split()[1].startswith(('DMUL', 'DADD', 'DFMA'))))
    st = collections.Counter({h[6:]: sum(I(r[hdr.index(h)]) for r in seg) for h in stalls})
    top = ", ".join("%s %.0f%%" % (k, 100 * v / max(s, 1)) for k, v in st.most_common(5))
    print("[%5d,%5d) time %5.1f%%  instr %10d  DP %10d  smem wavefronts %10d (excess %9d) | %s" % (a, b, 100 * s / max(tot, 1), e, dp, w, x, top))
