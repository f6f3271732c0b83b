"""Top stall instructions of an .ncu-rep captured with --import-source on: usage ncu_src.py rep [ntop]"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h0 = next(i for i, r in enumerate(rows) if r and r[0] == "Address")          # first kernel's table
hdr = rows[h0]
body = []
for r in rows[h0 + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    if len(r) >= len(hdr):
        body.append(r)
isamp, isrc, iexe = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[isamp] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
order = sorted(range(len(body)), key=lambda k: -int(body[k][isamp] or 0))[:ntop]
for k in sorted(order):
    r = body[k]
    top = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stalls), reverse=True)[:2]
    print(f"{k:4d} {int(r[isamp]):6d} {100 * int(r[isamp]) / tot:5.1f}%  exe={r[iexe]:>7s} {r[isrc].strip()[:70]:70s} {top}")
