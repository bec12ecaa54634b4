"""Extract per-launch DRAM traffic, duration and tensor-pipe activity from committed ncu reports into
profiles/r02_ncu_traffic.json -- bench.py reads `traffic` from this file instead of a hard-coded literal
(the literal disagreed with the committed capture).

   python tools/ncu_traffic.py name=profiles/file.ncu-rep[:kernel-substring] ...
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {"dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
        "gpu__time_duration.sum": "duration", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
        "launch__registers_per_thread": "registers", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
        "lts__t_bytes.sum": "l2_bytes", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct"}


def to_bytes(v, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(v) * m.get(unit, 1)


def parse(rep, match=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, data = rows[0], rows[1], rows[2:]
    res = []
    for r in data:
        d = dict(zip(head, r))
        name = d.get("Kernel Name", "")
        if match and match not in name:
            continue
        e = {"kernel": name, "grid": d.get("Grid Size"), "block": d.get("Block Size")}
        for k, nk in WANT.items():
            if k in d and d[k] != "":
                u = units[head.index(k)]
                v = d[k].replace(",", "")
                if "bytes" in nk:
                    e[nk] = to_bytes(v, u)
                elif nk == "duration":
                    e["duration_us"] = float(v) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(u, 1.0)
                else:
                    e[nk] = float(v)
        if "dram_read_bytes" in e and "dram_write_bytes" in e:
            e["traffic"] = e["dram_read_bytes"] + e["dram_write_bytes"]
        res.append(e)
    return res


def main():
    path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    try:
        cur = json.load(open(path))
    except Exception:
        cur = {}
    for arg in sys.argv[1:]:
        name, spec = arg.split("=", 1)
        rep, _, match = spec.partition(":")
        launches = parse(rep, match or None)
        if not launches:
            print("no launch matched", arg); continue
        e = launches[-1]
        cur[name] = e.get("traffic")
        cur[name + "__detail"] = dict(e, source=os.path.relpath(rep, ROOT), launches_in_report=len(launches))
        print(name, json.dumps(cur[name + "__detail"]))
    json.dump(cur, open(path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
