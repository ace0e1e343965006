"""Dev tool: registers / spills per kernel of one csrc/*.cu (nvcc -Xptxas -v, sm_100a).
    python tools/regs.py head_up.cu [regex on the mangled name]"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "restrictive-hierarchical-semantic-segmentation_b200")


def main(src, pat="."):
    cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
           "--expt-relaxed-constexpr", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(PKG, "csrc"), "-Xptxas=-v",
           "-c", os.path.join(PKG, "csrc", src), "-o", "/tmp/_regs_%s.o" % src]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        print(r.stderr[-4000:])
        sys.exit(1)
    lines = r.stderr.splitlines()
    cur = None
    for i, l in enumerate(lines):
        m = re.search(r"Compiling entry function '(\S+)'", l)
        if m:
            cur = m.group(1)
            continue
        if cur and "registers" in l and re.search(pat, cur):
            d = subprocess.run(["c++filt", cur], capture_output=True, text=True).stdout.strip()
            d = re.sub(r"\(.*", "", d).replace("void rhseg::", "")
            spill = ""
            for j in range(max(0, i - 3), i):
                if "spill" in lines[j]:
                    spill = lines[j].strip().replace("bytes ", "B ")
            print("%-70s %s regs | %s" % (d[:70], re.search(r"Used (\d+) registers", l).group(1), spill))
            cur = None


if __name__ == "__main__":
    main(*sys.argv[1:3])
