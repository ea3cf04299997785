"""ncu_extract.py — turns `ncu --set full` captures (gpurun_out/*.ncu-rep, not committed: MBs each) into what IS committed:

  profiles/<tag>.raw.csv                 `ncu -i <rep> --page raw --csv` transposed (metric, unit, value per captured launch),
                                         restricted to the metric families that matter for this path (a few kB)
  profiles/r02_kernel_counters.json      per workload key ("cfg2:levelset:tex"): warp instructions per launch, lanes per
                                         instruction, IPC, DRAM bytes, L1 / L2 hit rates, texture pipe — read by bench.py's roofline

usage: python tests/ncu_extract.py <key>=<file.ncu-rep>[@launch] ...      (runs here, no GPU needed)
"""
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = re.compile(r"^(Kernel Name|Block Size|Grid Size|gpu__time_duration|launch__(registers_per_thread|occupancy_limit|waves_per|shared_mem_per_block|grid_size|block_size)|"
                  r"sm__inst_executed\.(sum|avg\.per_cycle)|smsp__inst_executed\.sum$|smsp__thread_inst_executed\.sum$|smsp__thread_inst_executed_per_inst_executed|"
                  r"smsp__issue_active\.avg|smsp__inst_issued\.avg|sm__warps_active\.avg\.pct|smsp__warps_eligible\.avg|smsp__warps_active\.avg\.per|"
                  r"dram__bytes_(read|write)\.sum$|dram__throughput\.avg\.pct|gpu__dram_throughput|"
                  r"l1tex__t_sector_hit_rate|lts__t_sector_hit_rate|l1tex__throughput\.avg|lts__throughput\.avg|"
                  r"l1tex__data_pipe_tex_wavefronts\.avg|l1tex__texin_sm2tex_req_cycles_active\.avg|sm__inst_executed_pipe_(tex|alu|fma|fmaheavy|fmalite|lsu|xu|fp64|uniform|cbu|adu)\.avg\.pct_of_peak_sustained_active|"
                  r"sm__pipe_(alu|fma|fmaheavy|xu)_cycles_active\.avg\.pct_of_peak_sustained_active|"
                  r"sm__cycles_(active|elapsed)\.avg$|smsp__cycles_active\.avg$|sm__throughput\.avg\.pct|"
                  r"smsp__average_warps?_issue_stalled_[a-z_]+_per_issue_active|smsp__average_warp_latency_issue_stalled|"
                  r"sm__sass_thread_inst_executed_op_(fadd|fmul|ffma|integer)_pred_on\.sum$|smsp__sass_inst_executed_op_(texture|global_ld|shared|local)|"
                  r"l1tex__t_requests_pipe_tex_mem_texture\.sum$|l1tex__t_sectors_pipe_tex_mem_texture\.sum$|lts__t_sectors_srcunit_tex_op_read\.sum$)")


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def fnum(x):
    try:
        return float(x.replace(",", ""))
    except Exception:
        return None


def main():
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    jpath = os.path.join(ROOT, "profiles", "r02_kernel_counters.json")
    table = json.load(open(jpath)) if os.path.exists(jpath) else {}
    for arg in sys.argv[1:]:
        key, rep = arg.split("=", 1)
        pick = -1
        if "@" in rep:
            rep, n = rep.rsplit("@", 1)
            pick = int(n)
        hdr, units, data = raw_rows(rep)
        tag = os.path.splitext(os.path.basename(rep))[0]
        cols = [i for i, h in enumerate(hdr) if KEEP.match(h)]
        with open(os.path.join(ROOT, "profiles", tag + ".raw.csv"), "w", newline="") as f:
            wr = csv.writer(f)
            wr.writerow(["metric", "unit"] + [f"launch{k}" for k in range(len(data))])
            for i in cols:
                wr.writerow([hdr[i], units[i]] + [row[i] for row in data])
        row = data[pick]
        col = {h: i for i, h in enumerate(hdr)}

        def val(name, scale_units=True):
            i = col.get(name)
            if i is None:
                return None
            v = fnum(row[i])
            if v is None:
                return None
            u = units[i]
            if scale_units:
                v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(u, 1.0)   # bytes; durations in ms
            return v
        inst = val("smsp__inst_executed.sum")
        table[key] = {
            "kernel": row[col["Kernel Name"]], "source": f"profiles/{tag}.raw.csv (ncu --set full --clock-control none, launch {pick if pick >= 0 else len(data) - 1} of the capture)",
            "ncu_ms": val("gpu__time_duration.sum"), "registers": val("launch__registers_per_thread"),
            "inst_executed": inst, "thread_inst_executed": val("smsp__thread_inst_executed.sum"),
            "lanes_per_inst": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
            "ipc": val("sm__inst_executed.avg.per_cycle_elapsed"), "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "dram_bytes": (val("dram__bytes_read.sum") or 0) + (val("dram__bytes_write.sum") or 0),
            "dram_read_bytes": val("dram__bytes_read.sum"), "dram_write_bytes": val("dram__bytes_write.sum"),
            "l1_hit_pct": val("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": val("lts__t_sector_hit_rate.pct"),
            "tex_pipe_pct": val("l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "l1tex_throughput_pct": val("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        }
        print(key, json.dumps(table[key]))
    json.dump(table, open(jpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
