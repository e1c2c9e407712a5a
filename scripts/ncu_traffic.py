"""profiles/r2_traffic.json from an `ncu --set full` capture: DRAM bytes per launch (dram__bytes_read.sum +
dram__bytes_write.sum) of the dominant kernel, the number bench.py reports as roofline.traffic.

    python scripts/ncu_traffic.py gpurun_out/prof_r2_c3.ncu-rep c3 1 [more: report config n_gpus ...]
"""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def rows_of(report):
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def main():
    args = sys.argv[1:]
    captures = []
    for q in range(0, len(args), 3):
        report, config, n_gpus = args[q], args[q + 1], int(args[q + 2])
        hdr, units, rows = rows_of(report)
        kn, rd, wr, tm = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
        per_kernel = {}
        for r in rows:
            name = re.sub(r"\(.*", "", r[kn]).replace("void ", "")
            b = float(r[rd]) * UNIT[units[rd]] + float(r[wr]) * UNIT[units[wr]]
            per_kernel.setdefault(name, []).append((b, float(r[tm])))
        for name, v in per_kernel.items():
            captures.append({"kernel": name, "config": config, "n_gpus": n_gpus, "launches_captured": len(v),
                             "dram_bytes_per_launch": sum(b for b, _ in v) / len(v), "dram_bytes_each": [b for b, _ in v],
                             "ms_each_under_ncu": [t for _, t in v], "report": os.path.basename(report)})
    doc = {"source": "ncu --set full --clock-control none --import-source on; dram__bytes_read.sum + dram__bytes_write.sum per launch "
                     "(the .ncu-rep files stay in gpurun_out/, scratch); written by scripts/ncu_traffic.py", "captures": captures}
    with open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w") as fh:
        json.dump(doc, fh, indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
