"""profiles/r2_checkerboard_traffic.json from an `ncu --set full` capture of the bench's own launch of the sweep kernel:
usage traffic_json.py rep sweeps_per_launch method beta out.json"""
import csv
import json
import subprocess
import sys

rep, spl, method, beta, out = sys.argv[1], float(sys.argv[2]), sys.argv[3], float(sys.argv[4]), sys.argv[5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]


def col(name):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)


rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
json.dump({"kernel": r[hdr.index("Kernel Name")], "capture": rep, "sweeps_per_launch": spl, "method": method, "beta": beta,
           "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
           "note": "one launch of the bench's own command under ncu --set full --clock-control none --cache-control none "
                   "(the bench flushes L2 before the launch itself)"}, open(out, "w"), indent=1)
print(open(out).read())
