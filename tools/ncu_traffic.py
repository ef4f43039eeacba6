"""Turn `ncu --set full` reports into the table bench.py reads for `roofline.traffic` (profiles/ncu_traffic.json).

    python tools/ncu_traffic.py <report.ncu-rep> <workload> <states_per_gpu> <kernel-key> [<report> <workload> <states> <key> ...]

kernel-key is what bench.py looks up: "k_cg_solve" (persistent solve kernel), "fused_dmma" or "gemm_chain". For every kernel
instance in the report whose name contains the key's kernel name the DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum)
are averaged per launch; the raw numbers and the report name are kept beside them.

    python tools/ncu_traffic.py --chain-log <ncu --csv --log-file output> <workload> <states_per_gpu> <n_fvp>

GEMM-chain path: one FVP is a sequence of launches, so the figure is the SUM over every kernel in the log divided by the number of
FVPs the profiled command ran (`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum`); the per-kernel
split is kept in the entry."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MATCH = {"k_cg_solve": "k_cg_solve", "fused_dmma": "k_fvp_fused", "gemm_chain": "k_chain"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def read(report, key):
    raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        if MATCH[key] not in r[ix["Kernel Name"]]:
            continue
        rd = float(r[ix["dram__bytes_read.sum"]]) * UNIT[units[ix["dram__bytes_read.sum"]]]
        wr = float(r[ix["dram__bytes_write.sum"]]) * UNIT[units[ix["dram__bytes_write.sum"]]]
        out.append((rd, wr, float(r[ix["gpu__time_duration.sum"]]), units[ix["gpu__time_duration.sum"]], r[ix["Kernel Name"]][:100]))
    return out


def chain_log(log, workload, states, n_fvp):
    import collections
    import re
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    for r in csv.reader(open(log)):
        if len(r) > 14 and r[0].isdigit():
            name = re.sub(r"\(.*", "", r[4]).replace("void ", "").replace("<unnamed>::", "")
            per[name][r[12]] += float(r[14]) * (UNIT.get(r[13], 1) if r[12].startswith("dram") else 1)
            if r[12] == "gpu__time_duration.sum":
                per[name]["launches"] += 1
    rd = sum(d["dram__bytes_read.sum"] for d in per.values()) / n_fvp
    wr = sum(d["dram__bytes_write.sum"] for d in per.values()) / n_fvp
    return {"workload": workload, "states_per_gpu": states, "kernel": "gemm_chain", "dram_bytes_per_launch": rd + wr,
            "dram_bytes_read": rd, "dram_bytes_written": wr, "unit_of_a_launch": "one FVP (all kernels of the chain)",
            "fvps_in_log": n_fvp, "report": os.path.basename(log),
            "per_kernel_per_fvp": {k: {"launches": d["launches"] / n_fvp, "ms_under_ncu": d["gpu__time_duration.sum"] / n_fvp / 1e6,
                                       "dram_GB": (d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]) / n_fvp / 1e9}
                                   for k, d in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"])}}


def main():
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    table = json.load(open(path)) if os.path.exists(path) else {"entries": []}
    a = sys.argv[1:]
    if a and a[0] == "--chain-log":
        e = chain_log(a[1], a[2], int(a[3]), int(a[4]))
        table["entries"] = [x for x in table["entries"] if not (x["workload"] == e["workload"] and x["states_per_gpu"] == e["states_per_gpu"] and x["kernel"] == "gemm_chain")]
        table["entries"].append(e)
        print(json.dumps(e))
        json.dump(table, open(path, "w"), indent=1)
        return
    for i in range(0, len(a), 4):
        report, workload, states, key = a[i], a[i + 1], int(a[i + 2]), a[i + 3]
        inst = read(report, key)
        if not inst:
            sys.exit(f"no kernel matching {key} in {report}")
        rd = sum(x[0] for x in inst) / len(inst)
        wr = sum(x[1] for x in inst) / len(inst)
        e = {"workload": workload, "states_per_gpu": states, "kernel": key, "dram_bytes_per_launch": rd + wr,
             "dram_bytes_read": rd, "dram_bytes_written": wr, "launches_in_report": len(inst),
             "kernel_name": inst[0][4], "duration_under_ncu": f"{inst[0][2]} {inst[0][3]}", "report": os.path.basename(report)}
        table["entries"] = [x for x in table["entries"] if not (x["workload"] == workload and x["states_per_gpu"] == states and x["kernel"] == key)]
        table["entries"].append(e)
        print(json.dumps(e))
    json.dump(table, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
