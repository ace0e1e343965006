"""Dev tool: opcode tally of every kernel in librhseg_b200.so (cuobjdump -sass) -> profiles/rNN_sass_opcodes.txt.
    python tools/sass_tally.py [out.txt]
Per kernel: instruction count, the bulk-copy / mbarrier opcodes that prove the TMA-engine pipelines (UBLKCP = cp.async.bulk,
SYNCS = mbarrier ops, UTMALDG = tensor-map loads, UTC*MMA = tcgen05), the local-memory opcodes that prove spills (STL / LDL),
and the widths of shared / global accesses."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "restrictive-hierarchical-semantic-segmentation_b200", "librhseg_b200.so")
COLS = ["UBLKCP", "SYNCS", "UTMALDG", "UTCMMA", "STL", "LDL", "LDS.128", "LDS", "STG.128", "STG", "LDG.128", "LDG", "ATOM/RED", "SHFL", "FFMA", "DFMA/DADD"]


def classify(op):
    base = op.split(".")[0]
    out = []
    if base == "UBLKCP": out.append("UBLKCP")
    if base == "SYNCS": out.append("SYNCS")
    if base.startswith("UTMALDG"): out.append("UTMALDG")
    if base.startswith("UTC") and "MMA" in base: out.append("UTCMMA")
    if base in ("STL", "LDL"): out.append(base)
    if base in ("LDS", "STG", "LDG"):
        out.append(base + ".128" if ".128" in op else base)
    if base in ("ATOM", "ATOMG", "RED", "ATOMS"): out.append("ATOM/RED")
    if base == "SHFL": out.append("SHFL")
    if base == "FFMA": out.append("FFMA")
    if base in ("DFMA", "DADD", "DMUL"): out.append("DFMA/DADD")
    return out


def main(out_path=None):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    rows = []
    for blk in sass.split("Function : ")[1:]:
        name = blk.split("\n", 1)[0].strip()
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        dem = re.sub(r"\(.*", "", dem).replace("void rhseg::", "").replace("rhseg::", "")
        cnt = collections.Counter()
        n = 0
        for m in re.finditer(r"^\s+/\*[0-9a-f]{4,6}\*/\s+(.*?);", blk, flags=re.M):
            t = re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip())
            op = t.split()[0]
            n += 1
            for c in classify(op):
                cnt[c] += 1
        rows.append((dem, n, cnt))
    rows.sort(key=lambda r: r[0])
    lines = ["# cuobjdump -sass of %s (sm_100a): %d kernels" % (os.path.basename(LIB), len(rows))]
    tot = collections.Counter()
    for _, _, c in rows:
        tot.update(c)
    lines.append("# totals: " + "  ".join("%s=%d" % (k, tot[k]) for k in COLS))
    spill = [(d, c["STL"], c["LDL"]) for d, _, c in rows if c["STL"] or c["LDL"]]
    lines.append("# kernels with local-memory (spill) opcodes: %d of %d" % (len(spill), len(rows)))
    for d, a, b in spill:
        lines.append("#   STL=%-4d LDL=%-4d %s" % (a, b, d))
    lines.append("# kernels that stage through the bulk-copy engine (UBLKCP > 0): %d" % sum(1 for _, _, c in rows if c["UBLKCP"]))
    lines.append("")
    lines.append("%-84s %6s " % ("kernel", "instr") + " ".join("%8s" % c for c in COLS))
    for d, n, c in rows:
        lines.append("%-84s %6d " % (d[:84], n) + " ".join("%8d" % c[k] for k in COLS))
    text = "\n".join(lines) + "\n"
    if out_path:
        with open(out_path, "w") as f:
            f.write(text)
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main(*sys.argv[1:2])
