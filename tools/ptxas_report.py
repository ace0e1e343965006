"""Dev tool: registers / stack / spill bytes of EVERY kernel (nvcc -Xptxas=-v over csrc/*.cu, sm_100a) ->
profiles/rNN_ptxas_spills.txt.  A stack frame without spill bytes is a local array the kernel indexes at run time (by
design); spill stores / loads are registers ptxas could not keep.
    python tools/ptxas_report.py [out.txt]"""
import concurrent.futures
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "restrictive-hierarchical-semantic-segmentation_b200")


def one(src):
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "csrc"), "-Xptxas=-v",
           "-c", os.path.join(PKG, "csrc", src), "-o", "/tmp/_ptxas_%s.o" % src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise SystemExit(r.stderr[-3000:])
    rows, cur, frame = [], None, (0, 0, 0)
    for l in r.stderr.splitlines():
        m = re.search(r"Compiling entry function '(\S+)'", l)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", l)
        if m:
            frame = tuple(int(x) for x in m.groups())
            continue
        m = re.search(r"Used (\d+) registers", l)
        if m and cur:
            rows.append((src, cur, int(m.group(1))) + frame)
            cur, frame = None, (0, 0, 0)
    return rows


def main(out_path=None):
    srcs = sorted(f for f in os.listdir(os.path.join(PKG, "csrc")) if f.endswith(".cu"))
    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        rows = [r for rs in ex.map(one, srcs) for r in rs]
    names = subprocess.run(["c++filt"], input="\n".join(r[1] for r in rows), capture_output=True, text=True).stdout.splitlines()
    out = ["# nvcc -Xptxas=-v, sm_100a: %d kernels in %d translation units" % (len(rows), len(srcs))]
    spilled = [(r, n) for r, n in zip(rows, names) if r[4] or r[5]]
    out.append("# kernels with spill bytes: %d" % len(spilled))
    out.append("# %-12s %5s %6s %7s %7s  kernel" % ("file", "regs", "stack", "spill_st", "spill_ld"))
    for r, n in spilled:
        n = re.sub(r"\(.*", "", n).replace("void rhseg::", "").replace("rhseg::", "")
        out.append("  %-12s %5d %6d %7d %7d  %s" % (r[0], r[2], r[3], r[4], r[5], n))
    out.append("")
    out.append("# all kernels")
    for r, n in zip(rows, names):
        n = re.sub(r"\(.*", "", n).replace("void rhseg::", "").replace("rhseg::", "")
        out.append("  %-12s %5d %6d %7d %7d  %s" % (r[0], r[2], r[3], r[4], r[5], n))
    text = "\n".join(out) + "\n"
    if out_path:
        with open(out_path, "w") as f:
            f.write(text)
    print("\n".join(out[:len(spilled) + 4]))


if __name__ == "__main__":
    main(*sys.argv[1:2])
