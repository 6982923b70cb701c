#!/usr/bin/env python
"""Per-kernel summary of `ncu --set full` reports (read here with `ncu -i REP --page raw --csv`): DRAM bytes, duration, issue-slot
utilisation, instructions, registers.  usage: tools/ncu_traffic.py out.json name=report.ncu-rep [name=report.ncu-rep ...]
Merges into out.json's "kernels" map (profiles/r1_traffic.json is what bench.py reads `roofline.traffic` from)."""
import csv, io, json, os, subprocess, sys

WANT = {
    "dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
    "gpu__time_duration.sum": "duration_ms_under_ncu", "sm__inst_executed.sum": "inst_executed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.per_cycle_active": "warps_active_per_sm", "launch__registers_per_thread": "registers_per_thread",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflicts",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3,
        "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def main():
    out = sys.argv[1]
    doc = json.load(open(out)) if os.path.exists(out) else {"kernels": {}}
    for arg in sys.argv[2:]:
        name, rep = arg.split("=", 1)
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rows = list(csv.reader(io.StringIO(txt)))
        head, units, vals = rows[0], rows[1], rows[2]
        k = {}
        for h, u, v in zip(head, units, vals):
            if h in WANT and v not in ("", "n/a"):
                x = float(v.replace(",", ""))
                if WANT[h].endswith("bytes") or WANT[h].startswith("duration"):
                    x *= UNIT.get(u, 1.0)
                k[WANT[h]] = x
        k["kernel_name"] = vals[head.index("Kernel Name")] if "Kernel Name" in head else ""
        k["source"] = os.path.basename(rep)
        doc["kernels"][name] = k
        print(name, k)
    json.dump(doc, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
