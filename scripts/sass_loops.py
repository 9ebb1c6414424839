"""Static issue-slot census: for every kernel of a binary, the opcode counts of each backward-branch loop body.

    python scripts/sass_loops.py /tmp/tc_scm_ab            (any cubin / executable / .so that cuobjdump reads)

Used for profiles/tc_ab_r2.md (instructions per bin-frame of the packed-FP32 and the mma.sync covariance engines)."""
import collections
import re
import subprocess
import sys


def main(path, only=""):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n")[0]
        if only and only not in name:
            continue
        ins = [(int(m.group(1), 16), m.group(2).strip()) for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", f)]
        print("==", name, "(%d instructions)" % len(ins))
        for a, t in ins:
            mm = re.search(r"BRA.*?0x([0-9a-f]+)", t)
            if not mm or int(mm.group(1), 16) >= a:
                continue
            lo = int(mm.group(1), 16)
            body = [re.sub(r"^@!?U?P\d+\s+", "", x) for y, x in ins if lo <= y <= a]
            c = collections.Counter(x.split()[0].split(".")[0] for x in body)
            print("   loop %#06x..%#06x: %4d instructions  %s" % (lo, a, len(body), dict(c.most_common(12))))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
