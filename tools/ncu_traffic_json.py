#!/usr/bin/env python
"""profiles/ncu_traffic.json from the two traffic captures of tools/gpu_profile.sh:

    python tools/ncu_traffic_json.py <tag>      # reads profiles/<tag>_ncu_traffic_{batch4096,field4096}.csv

DRAM bytes (read + write) per launch of each of the four step kernels; bench.py reports the dominant kernel's figure as
`roofline.traffic`."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = {"gradient_forward_kernel": "physarum_forward", "move_claim_kernel": "move_claim",
         "field_step_kernel": "field_step", "agent_feed_kernel": "agent_feed"}


def per_kernel(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"')) ]
    hdr = rows[0]
    k, m, v = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    out, inst = {}, {}
    for r in rows[1:]:
        name = next((n for key, n in NAMES.items() if key in r[k]), None)
        if name is None or not r[m].startswith("dram__bytes"):
            continue
        out[name] = out.get(name, 0) + int(float(r[v].replace(",", "")))
        inst[name] = r[k].split("(")[0].replace("void ", "")
    return out, inst


def main():
    tag = sys.argv[1]
    res, insts = {}, {}
    for wl, suffix in (("physarum_batched_4096x256x256", "batch4096"), ("physarum_single_field_4096x4096", "field4096")):
        res[wl], insts[wl] = per_kernel(os.path.join(ROOT, "profiles", f"{tag}_ncu_traffic_{suffix}.csv"))
    res["_source"] = (f"profiles/{tag}_ncu_traffic_batch4096.csv and profiles/{tag}_ncu_traffic_field4096.csv (ncu --metrics "
                      "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, step 42 of each "
                      "workload; tools/gpu_profile.sh, tools/ncu_traffic_json.py); bytes per launch of the shipped default "
                      "kernels: " + json.dumps(insts))
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(res, f, indent=1)
    for wl in res:
        if wl != "_source":
            print(wl, res[wl])


if __name__ == "__main__":
    main()
