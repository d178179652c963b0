"""Turns the CSV exports of one `ncu --set full --import-source on` capture of bench.py into the two small files bench.py quotes:
  profiles/ncu_traffic.json   DRAM bytes per element of the volume kernel (roofline.traffic), tagged with the kernel-source hash
  profiles/fp64_flops.json    FP64 operations per DOF of each kernel of a stage, counted on the SASS (thread-level DADD/DMUL = 1,
                              DFMA = 2, DMMA = 512 per warp instruction)
usage: python scripts/make_profile_json.py <raw.csv> <source.csv> <nElem> <order> <key e.g. NS_standard_P7> <label> [fp64|hbm]"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import source_sha

raw, src, nElem, order, key, label = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], sys.argv[6]
bound = sys.argv[7] if len(sys.argv) > 7 else "hbm"
ndof = nElem * (order + 1) ** 3
num = lambda x: float(x.replace(",", "")) if x.strip() else 0.0

rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
traffic = {}
for r in rows[2:]:
    name = r[ki].split("<")[0].replace("void ", "").replace("h3d::", "")
    traffic.setdefault(name, []).append(num(r[ri]) * scale[units[ri]] + num(r[wi]) * scale[units[wi]])

rows = list(csv.reader(open(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
flops = {}
for a, b in zip(starts[:-1], starts[1:]):
    name = rows[a][1].split("<")[0].replace("void ", "").replace("h3d::", "")
    h = rows[a + 1]; data = [r for r in rows[a + 2:b] if len(r) > 5]
    si, ti, ei = h.index("Source"), h.index("Thread Instructions Executed"), h.index("Instructions Executed")
    if name in flops or not any("BAR" in r[si] or "LDG" in r[si] for r in data):
        continue
    f = 0.0
    for r in data:
        op = r[si].replace("@P0", "").replace("@!P0", "").split()
        op = [o for o in op if not o.startswith("@")]
        if not op:
            continue
        if op[0].startswith(("DADD", "DMUL")): f += num(r[ti])
        elif op[0].startswith("DFMA"): f += 2 * num(r[ti])
        elif op[0].startswith("DMMA"): f += 512 * num(r[ei])
    flops[name] = f / ndof

vol = [k for k in traffic if k.startswith("k_volume")][0]
tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
t = json.load(open(tj)) if os.path.exists(tj) else {}
t = {k: v for k, v in t.items() if k.startswith("P")}
t["P%d" % order] = {"k_volume_bytes_per_element": sum(traffic[vol]) / len(traffic[vol]) / nElem, "source": label, "source_sha": source_sha(),
                    "per_kernel_bytes_per_element": {k: sum(v) / len(v) / nElem for k, v in traffic.items()}}
json.dump(t, open(tj, "w"), indent=1)
fj = os.path.join(ROOT, "profiles", "fp64_flops.json")
f = json.load(open(fj)) if os.path.exists(fj) else {}
volf = [v for k, v in flops.items() if k.startswith("k_volume")][0]
f[key] = {"volume_flop_per_dof": volf, "stage_flop_per_dof": sum(flops.values()), "per_kernel_flop_per_dof": flops, "source": label, "bound": bound}
json.dump(f, open(fj, "w"), indent=1)
print(json.dumps(t["P%d" % order], indent=1)); print(json.dumps(f[key], indent=1))
