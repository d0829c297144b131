"""Static view of a kernel's SASS by source line (no GPU needed): compiles nothing, reads an object file.
   python tools/sass_loop.py OBJ KERNEL_SUBSTR                 per-line static instruction counts
   python tools/sass_loop.py OBJ KERNEL_SUBSTR LO HI [FILE]    the SASS (in address order) whose line info falls in
                                                               FILE:LO..HI (default file centre_impl.cuh), with labels
Development tool: the inner loops of the centre kernel are tuned against these listings."""
import collections
import os
import re
import subprocess
import sys
import tempfile

obj, kern = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3]) if len(sys.argv) > 3 else None
hi = int(sys.argv[4]) if len(sys.argv) > 4 else None
fname = sys.argv[5] if len(sys.argv) > 5 else "centre_impl.cuh"
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
inside = False
cur = ("?", 0)
per = collections.Counter()
out = []
for ln in dis:
    if ln.startswith("\t.section\t.text."):
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\.L_x_\d+:", ln):
        out.append((None, ln))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*);", ln)
    if m:
        per[cur] += 1
        out.append((cur, m.group(2).strip()))
if lo is None:
    tot = sum(per.values())
    print("total static instructions:", tot)
    for (f, l), n in sorted(per.items()):
        if n >= 6:
            print("%-20s %5d  %4d" % (f, l, n))
else:
    # address range from the first to the last instruction attributed to FILE:LO..HI, everything in between
    # (inlined helpers from other files included), labels kept
    idx = [k for k, (cur, txt) in enumerate(out) if cur is not None and cur[0] == fname and lo <= cur[1] <= hi]
    shown = 0
    for cur, txt in out[idx[0]:idx[-1] + 1]:
        if cur is None:
            print(txt)
        else:
            print("   %-14s %4d  %s" % (cur[0][:14], cur[1], txt))
            shown += 1
    print("instructions shown:", shown)
