"""Per-region (between CTA barriers) stall breakdown of an `ncu --page source --csv` export holding several kernels.
usage: python scripts/ncu_regions.py source.csv [kernel-substring] [--hot N]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
hot = int(sys.argv[sys.argv.index("--hot") + 1]) if "--hot" in sys.argv else 0
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
I = lambda x: int(float(x)) if x.strip().replace(".", "").isdigit() else 0
seen = set()
for a, b in zip(starts[:-1], starts[1:]):
    name = rows[a][1]
    hdr = rows[a + 1]; data = [r for r in rows[a + 2:b] if len(r) > 5]
    si = hdr.index("Source")
    if want not in name or not data or not any("BAR" in r[si] or "LDS" in r[si] for r in data): continue   # SASS sections only
    if name in seen: continue
    seen.add(name)
    ai = hdr.index("Warp Stall Sampling (All Samples)"); ei = hdr.index("Instructions Executed")
    wi = hdr.index("L1 Wavefronts Shared"); wx = hdr.index("L1 Wavefronts Shared Excessive")
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(I(r[ai]) for r in data)
    print("== %s  (%d SASS instructions, %d samples)" % (name[:110], len(data), tot))
    bars = [i for i, r in enumerate(data) if "BAR.SYNC" in r[si]]
    edges = [0] + bars + [len(data)]
    for x, y in zip(edges[:-1], edges[1:]):
        seg = data[x:y]
        s = sum(I(r[ai]) for r in seg); e = sum(I(r[ei]) for r in seg)
        w = sum(I(r[wi]) for r in seg); xs = sum(I(r[wx]) for r in seg)
        dp = sum(I(r[ei]) for r in seg if any(t in r[si] for t in ("DMUL", "DADD", "DFMA")))
        mm = sum(I(r[ei]) for r in seg if "DMMA" in r[si])
        st = collections.Counter({h[6:]: sum(I(r[hdr.index(h)]) for r in seg) for h in stalls})
        top = ", ".join("%s %.0f%%" % (k, 100 * v / max(s, 1)) for k, v in st.most_common(5))
        print("[%5d,%5d) time %5.1f%%  instr %10d  DP %10d DMMA %8d  smem wf %10d (excess %9d) | %s" % (x, y, 100 * s / max(tot, 1), e, dp, mm, w, xs, top))
    if hot:
        idx = sorted(range(len(data)), key=lambda i: -I(data[i][ai]))[:hot]
        for i in sorted(idx):
            r = data[i]
            st = collections.Counter({h[6:]: I(r[hdr.index(h)]) for h in stalls})
            print("   #%5d %5.2f%%  %-70s %s" % (i, 100 * I(r[ai]) / max(tot, 1), r[si][:70], ", ".join("%s %d" % kv for kv in st.most_common(2))))
