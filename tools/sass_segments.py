"""Split the SASS of the centre kernel at its block barriers and report, per segment, the share
of executed warp instructions, the FP64 share, stall samples and the dominant stall reasons
(reads an .ncu-rep captured with --import-source on).  usage: sass_segments.py rep [kernel-regex]"""
import csv, io, re, subprocess, sys, collections
csv.field_size_limit(10 ** 9)
rep = sys.argv[1]
kn = sys.argv[2] if len(sys.argv) > 2 else "k_centre"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kn],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ia = hdr.index("Source"); ie = hdr.index("Instructions Executed"); it = hdr.index("Thread Instructions Executed"); isamp = hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
seg = []; cur = None
def new(k): return {'n': 0, 'fp': 0, 's': 0, 'start': k, 'thr': 0, 'st': collections.Counter(), 'ops': collections.Counter()}
cur = new(0); tot = 0
for k, r in enumerate(rows[2:]):
    try: n = float(r[ie]); s = float(r[isamp]); t = float(r[it])
    except Exception: continue
    src = r[ia].strip()
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', src)
    op = m.group(2) if m else ''
    cur['n'] += n; cur['s'] += s; cur['thr'] += t; cur['ops'][op.split('.')[0]] += n
    for i, h in stall_cols:
        try: cur['st'][h] += float(r[i])
        except Exception: pass
    if op.startswith(('DFMA', 'DMUL', 'DADD', 'DSETP', 'MUFU')): cur['fp'] += n
    tot += n
    if op.startswith('BAR'):
        cur['end'] = k; seg.append(cur); cur = new(k + 1)
cur['end'] = -1; seg.append(cur)
ss = sum(x['s'] for x in seg) or 1
print("total %.1fM warp instr, %d samples" % (tot / 1e6, ss))
for x in seg:
    if x['n'] / tot > 0.004:
        top = ", ".join("%s %.0f%%" % (h[6:], 100 * v / max(sum(x['st'].values()), 1)) for h, v in x['st'].most_common(4))
        ops = ", ".join("%s %.0f%%" % (h, 100 * v / x['n']) for h, v in x['ops'].most_common(5))
        print("sass %5d..%5d instr %5.2f%% fp64 %4.1f%% samples %5.2f%% thr/instr %4.1f | %s | %s" % (
            x['start'], x['end'], 100 * x['n'] / tot, 100 * x['fp'] / max(x['n'], 1), 100 * x['s'] / ss, x['thr'] / max(x['n'], 1), top, ops))
