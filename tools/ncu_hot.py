"""Dev tool: hot spots of one kernel from an .ncu-rep source page (SASS view).
    python tools/ncu_hot.py report.ncu-rep kernel-regex [kernel-index] [top-n]
Prints: instructions executed by region (ranges between the biggest-count changes), the top stalled instructions with
their dominant stall reason, and a stall-reason total."""
import csv
import io
import subprocess
import sys


def main(path, kre, which="0", top="40"):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                         capture_output=True, text=True).stdout
    # the CSV holds one block per kernel instance: split on the "Kernel Name" lines
    blocks, cur = [], None
    for line in out.splitlines():
        if line.startswith('"Kernel Name"'):
            cur = [line]
            blocks.append(cur)
        elif cur is not None:
            cur.append(line)
    blk = blocks[int(which)]
    print(blk[0][:160])
    rows = list(csv.reader(io.StringIO("\n".join(blk[1:]))))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    data = rows[1:]
    tot_inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
    tot_samp = sum(int(r[ix["# Samples"]]) for r in data)
    print("instructions executed (warp): %d   samples: %d   static SASS: %d" % (tot_inst, tot_samp, len(data)))
    totals = {c: sum(int(r[ix[c]]) for r in data) for c in stall_cols}
    print("stall totals: " + "  ".join("%s=%.1f%%" % (c[6:], 100.0 * v / max(1, tot_samp)) for c, v in sorted(totals.items(), key=lambda kv: -kv[1])[:9]))
    # opcode histogram weighted by executions
    ops = {}
    for r in data:
        src = r[ix["Source"]].strip()
        parts = src.split()
        op = parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "?")
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + int(r[ix["Instructions Executed"]])
    print("executed by opcode: " + "  ".join("%s=%.1f%%" % (k, 100.0 * v / tot_inst) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:22]))
    print("top stalled instructions:")
    order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:int(top)]
    for i in sorted(order):
        r = data[i]
        s = int(r[ix["# Samples"]])
        dom = max(stall_cols, key=lambda c: int(r[ix[c]]))
        print("  #%-5d %5.2f%%  exec=%-8s %-12s %s" % (i, 100.0 * s / max(1, tot_samp), r[ix["Instructions Executed"]], dom[6:], r[ix["Source"]].strip()[:90]))


if __name__ == "__main__":
    main(*sys.argv[1:5])
