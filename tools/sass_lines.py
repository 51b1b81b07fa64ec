"""Static SASS of a kernel by source line range (nvdisasm --print-line-info of the cubin).

    cuobjdump -xelf all hsr_env_b200/csrc/hsrb_wpe.o; nvdisasm --print-line-info hsrb_wpe.sm_100a.cubin > /tmp/wpe_lines.sass
    python tools/sass_lines.py /tmp/wpe_lines.sass _Z17hsrb_wpe_kernel_tILb1EEv5KArgs8PushInfo hsrb_wpe.cuh 785 850 [--dump]
"""
import collections
import re
import sys

path, fn, fname, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
dump = "--dump" in sys.argv
inside = False
cur = (None, 0)
ops = collections.Counter()
lines = collections.Counter()
n = 0
for l in open(path):
    if l.startswith(".text."):
        inside = l.strip() == f".text.{fn}:"
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m and cur[0] == fname and lo <= cur[1] <= hi:
        n += 1
        ops[m.group(2).split(".")[0]] += 1
        lines[cur[1]] += 1
        if dump:
            print(cur[1], l.rstrip()[:140])
print(f"{n} static SASS instructions for {fname}:{lo}-{hi}")
print(ops.most_common(25))
print(sorted(lines.items()))
