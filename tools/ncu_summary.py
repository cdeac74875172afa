"""Summarise an .ncu-rep (first kernel) into text: key sections, top stall reasons, DRAM traffic.
Usage: python tools/ncu_summary.py report.ncu-rep [--json profiles/<name>.json --shape N,M,T,d] > profiles/<name>.txt
The JSON (kernel name, launch shape, DRAM bytes per launch, duration) is what bench.py reads for `roofline.traffic`."""
import csv
import io
import json
import subprocess
import sys


def ncu(args):
    return subprocess.run(["ncu", "-i"] + args, capture_output=True, text=True).stdout


rep = sys.argv[1]
rows = list(csv.reader(io.StringIO(ncu([rep, "--page", "details", "--csv"]))))
hdr = rows[0]
isec, iname, iu, iv, ik = (hdr.index(k) for k in ("Section Name", "Metric Name", "Metric Unit", "Metric Value", "Kernel Name"))
print("# kernel:", rows[1][ik])
keep = ["GPU Speed Of Light Throughput", "Compute Workload Analysis", "Memory Workload Analysis", "Scheduler Statistics",
        "Warp State Statistics", "Launch Statistics", "Occupancy"]
skip = ("Cluster", "Green", "TPC", "Stack", "Function Cache", "Compression", "Driver Shared")
for r in rows[1:]:
    if r[isec] in keep and not any(s in r[iname] for s in skip) and r[iname]:
        print(f"{r[isec][:26]:26s} | {r[iname][:44]:44s} | {r[iu]:16s} | {r[iv]}")
raw = list(csv.reader(io.StringIO(ncu([rep, "--page", "raw", "--csv"]))))
h, u, v = raw[0], raw[1], raw[2]
print("# raw metrics")
for i, name in enumerate(h):
    if name in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
                "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
                "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
                "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
                "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"):
        print(f"{name} = {v[i]} {u[i]}")
st = []
for i, name in enumerate(h):
    if "pcsamp_warps_issue_stalled" in name and "not_issued" not in name:
        try:
            st.append((float(v[i].replace(",", "")), name.replace("smsp__pcsamp_warps_issue_stalled_", "")))
        except ValueError:
            pass
tot = sum(x for x, _ in st) or 1.0
print("# warp stall samples: " + ", ".join(f"{n} {100 * x / tot:.1f}%" for x, n in sorted(st, reverse=True)[:7]))

if "--json" in sys.argv:
    def num(name):
        i = h.index(name)
        x = float(v[i].replace(",", ""))
        unit = u[i].lower()
        scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1,
                 "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1}.get(unit, 1)
        return x * scale

    rec = {"kernel": rows[1][ik].replace("void ", "").replace("socks::", ""), "report": rep,
           "dram_bytes_read": num("dram__bytes_read.sum"), "dram_bytes_write": num("dram__bytes_write.sum"),
           "duration_s": num("gpu__time_duration.sum")}
    if "--shape" in sys.argv:
        rec["shape"] = [int(x) for x in sys.argv[sys.argv.index("--shape") + 1].split(",")]
    for name in ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                 "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
                 "launch__registers_per_thread"):
        if name in h:
            rec[name] = float(v[h.index(name)].replace(",", ""))
    json.dump(rec, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
