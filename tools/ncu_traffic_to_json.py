"""ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch) of the contraction launches
of one train step -> profiles/contraction_dram_traffic.json (read by bench.py for roofline.traffic).

    python tools/ncu_traffic_to_json.py <csv> "<command that produced it>" """
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, source = sys.argv[1], sys.argv[2]
    rows = []
    with open(path) as fh:
        lines = [l for l in fh if l.startswith('"')]
    per = {}
    for r in csv.DictReader(lines):
        per.setdefault(r["ID"], {"kernel": r["Kernel Name"][:60]})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    rd = sum(v.get("dram__bytes_read.sum", 0.0) for v in per.values())
    wr = sum(v.get("dram__bytes_write.sum", 0.0) for v in per.values())
    out = {"bytes_per_step": rd + wr, "read_bytes": rd, "write_bytes": wr, "launches": len(per),
           "source": source + " (%s): sum over the %d contraction launches of one train step" % (
               os.path.relpath(path, ROOT), len(per))}
    with open(os.path.join(ROOT, "profiles", "contraction_dram_traffic.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
