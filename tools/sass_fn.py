"""Print the SASS of one kernel of a built library (substring match on the mangled name) and
an opcode histogram:  python tools/sass_fn.py movis_b200/csrc/libmovis_b200.so k_stack_opaque [--hist]"""
import re
import subprocess
import sys
from collections import Counter

lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, keep = None, []
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur and pat in cur and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
        keep.append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", line.strip()))
if "--hist" in sys.argv:
    c = Counter()
    for l in keep:
        toks = l.split()
        op = toks[1] if not toks[1].startswith("@") else toks[2]
        c[op.rstrip(";")] += 1
    print(len(keep), "instructions")
    for op, n in c.most_common(60):
        print(f"{n:6d} {op}")
else:
    print("\n".join(keep))
