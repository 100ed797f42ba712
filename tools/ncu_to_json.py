#!/usr/bin/env python
"""Machine-readable summary of one kernel of an .ncu-rep (what bench.py reads for roofline.traffic / executed FP64).
usage: python tools/ncu_to_json.py rep.ncu-rep out.json kernel-name workload pairs_per_launch "capture command" """
import csv, io, json, subprocess, sys

rep, out, kernel, workload, ppl, cmd = sys.argv[1:7]
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
f = lambda k: float(d[k].replace(",", ""))
scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
rd = f("dram__bytes_read.sum") * scale[u["dram__bytes_read.sum"]]
wr = f("dram__bytes_write.sum") * scale[u["dram__bytes_write.sum"]]
dur = f("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6}[u["gpu__time_duration.sum"]]
cyc = f("sm__cycles_elapsed.max")
pc = {k: f("smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % k) for k in ("dadd", "dmul", "dfma")}
res = {
    "source": cmd + "; read with ncu -i ... --page raw --csv (tools/ncu_to_json.py)", "kernel": kernel, "workload": workload,
    "pairs_per_launch": int(ppl), "duration_ms": dur, "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes": rd + wr,
    "warp_instructions": f("smsp__inst_executed.sum"), "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "fp64_pipe_pct": f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"), "fp64_thread_inst_per_cycle": pc,
    "sm_cycles_elapsed": cyc, "registers_per_thread": int(f("launch__registers_per_thread")),
    "spill_instructions": (f("sass__inst_executed_register_spilling") if d.get("sass__inst_executed_register_spilling", "") not in ("", "n/a", "no data") else 0.0),
    "stalls_per_issue": {k.split("stalled_")[1].split("_per_issue")[0]: float(v) for k, v in d.items()
                         if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
                         and v not in ("", "no data") and float(v) > 0.05},
    "executed_fp64_flop": (pc["dadd"] + pc["dmul"] + 2 * pc["dfma"]) * cyc,
}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res, indent=1))
