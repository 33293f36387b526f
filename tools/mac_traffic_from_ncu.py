#!/usr/bin/env python
"""Regenerate profiles/mac_traffic.json (what bench.py reports as roofline.traffic) from an `ncu --set full`
raw CSV of the MAC launches of ONE tiered period:
    ncu -i gpurun_out/prof_macp.ncu-rep --page raw --csv > profiles/r02_mac_tiers_ncu_raw.csv
    python tools/mac_traffic_from_ncu.py profiles/r02_mac_tiers_ncu_raw.csv 4096
Asserts that the captured kernels are the MAC kernels (k_mac_p / k_mac) and that there are as many launches as
tiers, so a stale or mismatched capture cannot silently feed the bench line."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, instances = sys.argv[1], int(sys.argv[2])
    n_tiers = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    rows = list(csv.reader(open(path)))
    hdr, data = rows[0], [r for r in rows[2:] if len(r) == len(rows[0])]
    col = {h: i for i, h in enumerate(hdr)}
    units = rows[1]

    def to_bytes(r, name):
        v, u = float(r[col[name]]), units[col[name]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]

    launches = []
    for r in data:
        name = r[col["Kernel Name"]]
        assert "k_mac" in name, f"not a MAC kernel: {name}"
        launches.append({"kernel": name[:80], "us": float(r[col["gpu__time_duration.sum"]]),
                         "dram_read": to_bytes(r, "dram__bytes_read.sum"), "dram_write": to_bytes(r, "dram__bytes_write.sum")})
    assert len(launches) == n_tiers, f"expected {n_tiers} MAC launches of one period, got {len(launches)}"
    total = sum(x["dram_read"] + x["dram_write"] for x in launches)
    out_path = os.path.join(ROOT, "profiles", "mac_traffic.json")
    try:
        j = json.load(open(out_path))
    except Exception:
        j = {}
    j["tiered"] = {"source": os.path.relpath(os.path.abspath(path), ROOT), "instances": instances, "launches": launches,
                   "dram_bytes_per_period": total, "dram_bytes_per_instance_period": total / instances,
                   "generated_by": "tools/mac_traffic_from_ncu.py"}
    json.dump(j, open(out_path, "w"), indent=1)
    print(json.dumps(j["tiered"], indent=1))


if __name__ == "__main__":
    main()
