"""SASS listings of the headline kernels (instruction text only) + opcode histogram.
usage: python scripts/sass_listing.py <lib.so> <outdir> [n]"""
import collections, os, re, subprocess, sys
lib, out = sys.argv[1], sys.argv[2]
n = sys.argv[3] if len(sys.argv) > 3 else "8"
os.makedirs(out, exist_ok=True)
txt = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)[1:]
want = {"k_gradientILi%sELb1ELb0ELb0" % n: "k_gradient_tma", "k_volumeILi%sELi0ELb1ELb0ELb0" % n: "k_volume_tma", "k_volumeILi%sELi1ELb1" % n: "k_volume_split_tma",
        "k_gradientILi%sELb1ELb0ELb1" % n: "k_gradient_tma_dmma", "k_volumeILi%sELi0ELb1ELb0ELb1" % n: "k_volume_tma_dmma",
        "k_volume2ILb0": "k_volume2", "k_volume2ILb1": "k_volume2_dmma", "k_gradient2ILb0": "k_gradient2", "k_gradient2ILb1": "k_gradient2_dmma",
        "k_face_h_pack": "k_face_h_pack", "k_face_h_min": "k_face_h_min",
        "k_riemannILi%sELb0" % n: "k_riemann", "k_riemannILi%sELb1" % n: "k_riemann_ext", "k_prolong_qILi%s" % n: "k_prolong_q", "k_red_residual": "k_red_residual",
        "k_red_timestep": "k_red_timestep", "k_red_integralsE": "k_red_integrals", "k_red_integrals2": "k_red_integrals2", "k_red_ke_balance": "k_red_ke_balance",
        "k_stage_limiter": "k_stage_limiter", "k_halo_pack": "k_halo_pack", "k_halo_unpack": "k_halo_unpack",
        "k_gradientILi%sELb0ELb1" % n: "k_gradient_general", "k_volumeILi%sELi2ELb0ELb0" % n: "k_volume_split_ext", "k_volumeILi%sELi0ELb0ELb1" % n: "k_volume_gradvars",
        "k_aos_to_soa_range": "k_aos_to_soa_range", "k_soa_to_aos_range": "k_soa_to_aos_range"}
summary = []
for f in funcs:
    name = f.split("\n", 1)[0].strip()
    for key, short in want.items():
        if key in name:
            ins = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(.*?);", f)
            ops = collections.Counter((i.split()[1] if i.startswith("@") else i.split()[0]).split(".")[0] for i in ins if i.strip())
            with open(os.path.join(out, "sass_%s_n%s.txt" % (short, n)), "w") as fh:
                fh.write("// %s\n// %d instructions; opcode histogram: %s\n" % (name, len(ins), dict(ops.most_common(25))))
                fh.write("\n".join(ins) + "\n")
            summary.append("%-22s %6d instr  DMMA %3d DFMA %4d DMUL %4d DADD %4d LDS %4d STS %4d LDG %4d STG %4d UBLKCP %3d SYNCS %3d BAR %2d" % (
                short, len(ins), ops["DMMA"], ops["DFMA"], ops["DMUL"], ops["DADD"], ops["LDS"], ops["STS"], ops["LDG"], ops["STG"], ops["UBLKCP"], ops["SYNCS"], ops["BAR"]))
open(os.path.join(out, "sass_summary.txt"), "w").write("\n".join(sorted(summary)) + "\n")
print("\n".join(sorted(summary)))
