"""Summarise an `ncu --page source --csv` dump: stall samples between BAR.SYNC landmarks + top instructions."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[idx['# Samples']] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
phase, acc, accs, execs, wf, wfx = 0, 0, {}, 0, 0, 0
def flush(label):
    global acc, accs, execs, wf, wfx
    top = sorted(accs.items(), key=lambda x: -x[1])[:4]
    print("phase %-28s samples %6d (%5.1f%%) warp-inst %9d smem-wf %9d (excess %9d) top %s" % (label, acc, 100.0 * acc / tot, execs, wf, wfx, [(k[6:], v) for k, v in top]))
    acc, accs, execs, wf, wfx = 0, {}, 0, 0, 0
for i, r in enumerate(data):
    sm = int(r[idx['# Samples']] or 0)
    acc += sm
    execs += int(r[idx['Instructions Executed']] or 0)
    wf += int(float(r[idx.get("L1 Wavefronts Shared", 0)] or 0)) if "L1 Wavefronts Shared" in idx else 0; wfx += int(float(r[idx["L1 Wavefronts Shared Excessive"]] or 0)) if "L1 Wavefronts Shared Excessive" in idx else 0
    for h in stalls:
        accs[h] = accs.get(h, 0) + int(r[idx[h]] or 0)
    if 'BAR.SYNC' in r[idx['Source']] or 'EXIT' in r[idx['Source']]:
        flush("up to #%d %s" % (i, r[idx['Source']].strip()[:18]))
flush("tail")
print("--- top instructions")
order = sorted(range(len(data)), key=lambda i: -int(data[i][idx['# Samples']] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]
for i in sorted(order):
    r = data[i]; sm = int(r[idx['# Samples']] or 0)
    top = sorted(((int(r[idx[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print(i, r[idx['Source']][:60].ljust(60), sm, "%.1f%%" % (100.0 * sm / tot), top, 'exec', r[idx['Instructions Executed']])
