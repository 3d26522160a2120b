"""Summarise one kernel of an ncu report into the JSON kept under profiles/.

    ncu -i REPORT.ncu-rep --page raw --csv > raw.csv
    python profiles/summarise_ncu.py raw.csv "description" > profiles/rNN_vK_render_kernel_full.json

Also prints (stderr) the traffic block bench.py reads from profiles/traffic.json.
"""
import csv
import json
import re
import sys

KEEP = re.compile(
    r"^(dram__bytes_(read|write)\.sum|gpu__time_duration\.sum|l1tex__data_pipe_lsu_wavefronts\.avg\.pct|"
    r"l1tex__t_output_wavefronts_pipe_lsu_mem_(global|local)_op_(ld|st)\.sum|l1tex__t_sector_hit_rate|"
    r"l1tex__throughput\.avg\.pct|launch__(grid_size|registers_per_thread|shared_mem_config_size|shared_mem_per_block_static)|"
    r"lts__t_sector_hit_rate|lts__throughput\.avg\.pct|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__issue_active\.avg\.pct|sm__pipe_(alu|fma)_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"sass__inst_executed_(register_spilling|local_loads|local_stores)|"
    r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|smsp__inst_executed\.sum|"
    r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__thread_inst_executed_per_inst_executed\.ratio)")


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, zip(vals, units)))
    out = {"kernel": sys.argv[2] if len(sys.argv) > 2 else d["Kernel Name"][0], "kernel_name": d["Kernel Name"][0],
           "metrics": {k: list(v) for k, v in sorted(d.items()) if KEEP.match(k)}}
    print(json.dumps(out, indent=1))

    def num(k):
        v, u = d[k]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        return float(v) * scale

    traffic = {
        "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
        "kernel_ms_under_ncu": float(d["gpu__time_duration.sum"][0]),
        "issue_active_pct": float(d["sm__issue_active.avg.pct_of_peak_sustained_elapsed"][0]),
        "threads_per_warp_instruction": float(d["smsp__thread_inst_executed_per_inst_executed.ratio"][0]),
        "warp_instructions": float(d["smsp__inst_executed.sum"][0]),
        "l1_hit_pct": float(d["l1tex__t_sector_hit_rate.pct"][0]),
        "l2_hit_pct": float(d["lts__t_sector_hit_rate.pct"][0]),
        "l1_throughput_pct": float(d["l1tex__throughput.avg.pct_of_peak_sustained_elapsed"][0]),
        "warps_active_pct": float(d["sm__warps_active.avg.pct_of_peak_sustained_active"][0]),
    }
    print(json.dumps(traffic, indent=1), file=sys.stderr)


if __name__ == "__main__":
    main()
