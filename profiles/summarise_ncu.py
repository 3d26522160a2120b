"""Summarise one kernel of an ncu report into the JSON kept under profiles/.

    ncu -i REPORT.ncu-rep --page raw --csv > raw.csv
    python profiles/summarise_ncu.py raw.csv "description" [key] > profiles/rNN_<what>_full.json

Prints the full metric selection on stdout.  With a third argument `key` (c3, c4, modeA,
...) it also merges the headline block -- DRAM bytes per launch and the utilisation of every
unit that could bind, with the busiest one named -- into profiles/binding.json, which
bench.py attaches to its `roofline` object for the matching workload at N = 1.
"""
import csv
import json
import re
import sys
from pathlib import Path

KEEP = re.compile(
    r"^(dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__time_duration\.sum|"
    r"l1tex__data_pipe_lsu_wavefronts\.avg\.pct_of_peak_sustained_elapsed|"
    r"l1tex__t_output_wavefronts_pipe_lsu_mem_(global|local)_op_(ld|st)\.sum|l1tex__t_sector_hit_rate\.pct|"
    r"l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|launch__(grid_size|registers_per_thread|shared_mem_config_size|shared_mem_per_block_static)|"
    r"lts__t_sector_hit_rate\.pct|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|lts__t_bytes\.sum|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__issue_active\.avg\.pct_of_peak_sustained_elapsed|sm__inst_executed_pipe_(alu|fma|xu|lsu)\.sum|"
    r"sm__pipe_(alu|fma|fmaheavy|xu)_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"sass__inst_executed_(register_spilling|local_loads|local_stores)|"
    r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|smsp__inst_executed\.sum|"
    r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__thread_inst_executed_per_inst_executed\.ratio)$")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    # one row per captured launch: take the last one (the warmed-up launch) unless told otherwise
    vals = rows[-1] if len(rows) > 2 else rows[2]
    d = dict(zip(hdr, zip(vals, units)))
    desc = sys.argv[2] if len(sys.argv) > 2 else d["Kernel Name"][0]
    out = {"kernel": desc, "kernel_name": d["Kernel Name"][0],
           "metrics": {k: list(v) for k, v in sorted(d.items()) if KEEP.match(k)}}
    print(json.dumps(out, indent=1))

    def num(k, default=0.0):
        if k not in d or d[k][0] == "":
            return default
        v, u = d[k]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
        return float(v) * scale

    ms = num("gpu__time_duration.sum")
    if d["gpu__time_duration.sum"][1] == "us":
        ms /= 1e3
    elif d["gpu__time_duration.sum"][1] == "ns":
        ms /= 1e6
    elif d["gpu__time_duration.sum"][1] in ("s", "second"):
        ms *= 1e3
    issue = num("sm__issue_active.avg.pct_of_peak_sustained_elapsed")
    lanes = num("smsp__thread_inst_executed_per_inst_executed.ratio")
    units_pct = {
        "sm_issue_slots": issue,
        "l1_data_pipe": num("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "l2": num("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        "dram": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    }
    stalls = {k.split("issue_stalled_")[1].split("_per_issue")[0]: float(v[0]) for k, v in d.items()
              if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v[0]}
    top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
    dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
    block = {
        "what": desc,
        "kernel_ms_under_ncu": round(ms, 4),
        "dram_bytes_per_launch": dram,
        "dram_GBps": round(dram / (ms * 1e-3) / 1e9, 1) if ms else None,
        "unit_busy_pct": {k: round(v, 2) for k, v in units_pct.items()},
        "busiest_unit": max(units_pct, key=units_pct.get),
        "threads_per_warp_instruction": lanes,
        "lane_issue_efficiency": round(issue / 100.0 * lanes / 32.0, 4),
        "warps_active_pct": round(num("sm__warps_active.avg.pct_of_peak_sustained_active"), 2),
        "l1_hit_pct": round(num("l1tex__t_sector_hit_rate.pct"), 2),
        "l2_hit_pct": round(num("lts__t_sector_hit_rate.pct"), 2),
        "warp_instructions": num("smsp__inst_executed.sum"),
        "registers_per_thread": num("launch__registers_per_thread"),
        "spill_instructions": num("sass__inst_executed_register_spilling"),
        "top_stalls_per_issue": {k: round(v, 2) for k, v in top},
    }
    print(json.dumps(block, indent=1), file=sys.stderr)
    if len(sys.argv) > 3:
        p = Path(__file__).resolve().parent / "binding.json"
        allb = json.loads(p.read_text()) if p.exists() else {}
        block["source"] = sys.argv[4] if len(sys.argv) > 4 else ""
        allb[sys.argv[3]] = block
        p.write_text(json.dumps(allb, indent=1) + "\n")


if __name__ == "__main__":
    main()
