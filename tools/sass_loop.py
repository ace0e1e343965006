"""Dev tool: opcode histogram of one kernel's SASS (optionally between two addresses).
    python tools/sass_loop.py obj 'mangled-substring' [lo hi]"""
import collections
import re
import subprocess
import sys


def main(obj, pat, lo=None, hi=None):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    blocks = out.split("Function : ")
    for blk in blocks[1:]:
        name = blk.split("\n", 1)[0].strip()
        if not re.search(pat, name):
            continue
        ins = re.findall(r"^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", blk, flags=re.M)
        lo_i = int(lo, 16) if lo else 0
        hi_i = int(hi, 16) if hi else 1 << 30
        sel = [(int(a, 16), t) for a, t in ins if lo_i <= int(a, 16) <= hi_i]
        hist = collections.Counter()
        for _, t in sel:
            t = re.sub(r"^@!?U?P\d+\s+", "", t.strip())
            hist[t.split()[0].split(".")[0]] += 1
        print(subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:100])
        print("  %d instructions%s" % (len(sel), "" if lo is None else " in [%s, %s]" % (lo, hi)))
        print("  " + "  ".join("%s %d" % kv for kv in hist.most_common(24)))
        # backward branches = loops
        for a, t in sel:
            m = re.search(r"BRA\S*\s+(?:\S+,\s+)?(0x[0-9a-f]+)", t)
            if m and int(m.group(1), 16) < a:
                print("  loop: 0x%x -> %s (%d instr)" % (a, m.group(1), (a - int(m.group(1), 16)) // 16 + 1))


if __name__ == "__main__":
    main(*sys.argv[1:5])
