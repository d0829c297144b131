"""Where a kernel spills: source lines of the LDL / STL instructions in an object file (static view).
   python tools/spill_lines.py OBJ KERNEL_SUBSTR"""
import os, re, subprocess, sys, tempfile
obj, kern = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
inside = False; cur = None
for ln in dis:
    if ln.startswith("\t.section\t.text."):
        inside = kern in ln; continue
    if not inside: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.search(r'\b(LDL|STL)\b', ln): print("%s:%d  %s" % (cur[0], cur[1], ln.strip()[12:80]))
