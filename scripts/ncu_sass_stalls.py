"""Per-instruction stall samples of one kernel from `ncu -i X.ncu-rep --page source --csv --print-source sass`.
   python scripts/ncu_sass_stalls.py source.csv [kernel index] [from offset hex] [to offset hex]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 30
secs, sec, hdr, name = [], [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        if sec:
            secs.append((name, hdr, sec))
        name, sec, hdr = r[1], [], None
    elif r and r[0] == "Address":
        hdr = r
    elif hdr:
        sec.append(r)
secs.append((name, hdr, sec))
name, hdr, sec = secs[k]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print(name, len(sec), "instructions; total samples", sum(int(r[ix["# Samples"]]) for r in sec))
base = int(sec[0][0], 16)
acc = 0
for r in sec:
    off = int(r[0], 16) - base
    if lo <= off <= hi:
        s = int(r[ix["# Samples"]])
        acc += s
        top = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
        print("%04x %-62s %6d ex=%7s %s" % (off, r[1].strip()[:62], s, r[ix["Instructions Executed"]],
                                             " ".join("%s=%d" % (h, v) for v, h in top if v > 0)))
print("samples in range", acc)
