#!/usr/bin/env python
"""Condense `ncu --page raw --csv` output into the per-kernel numbers bench.py and DESIGN.md cite.

usage: python profiles/summarize_ncu.py gpurun_out/tick_raw.csv profiles/r1_tick_kernels.json [launch_list.csv]
"""
import csv
import json
import sys

WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_read_sectors",
    "lts__t_sectors_srcunit_tex_op_write.sum": "l2_write_sectors",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "smsp__inst_executed.sum": "warp_instructions",
}
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}


def main(raw, out):
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        k = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("<unnamed>::", "").replace("void ", "").split("<")[0]}
        for metric, name in WANT.items():
            if metric in hdr:
                i = hdr.index(metric)
                v = float(r[i].replace(",", ""))
                k[name] = v * SCALE.get(units[i], 1.0)
        k["dram_bytes"] = k.get("dram_read", 0) + k.get("dram_write", 0)
        if k.get("duration"):
            k["dram_gbs"] = k["dram_bytes"] / k["duration"] / 1e9
        kernels.append(k)
    total = sum(k["duration"] for k in kernels)
    for k in kernels:
        k["share_of_profiled_time"] = k["duration"] / total
    json.dump({"source": raw, "kernels": kernels, "tick_dram_bytes": sum(k["dram_bytes"] for k in kernels),
               "tick_duration_s_under_ncu": total}, open(out, "w"), indent=1)
    for k in kernels:
        print(f'{k["kernel"]:12s} {k["duration"]*1e6:8.1f} us  {k["dram_bytes"]/1e6:8.1f} MB  {k.get("dram_gbs", 0):7.0f} GB/s  '
              f'share {k["share_of_profiled_time"]:.2f}')


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
