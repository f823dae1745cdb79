"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into profiles/<name>.json + a text table.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_buoyan_dilute [ncols]"""
import csv, io, json, subprocess, sys

rep, outbase = sys.argv[1], sys.argv[2]
ncols = int(sys.argv[3]) if len(sys.argv) > 3 else 55296
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "lts__t_sector_hit_rate.pct"]
out = []
for r in rows[2:]:
    d = {}
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            v = r[i]
            try:
                v = float(v.replace(",", ""))
            except ValueError:
                pass
            d[k] = {"value": v, "unit": units[i]}
    stalls = {}
    for i, h in enumerate(hdr):
        if "average_warps_issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h:
            stalls[h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")] = float(r[i])
    d["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
    # per-cycle sums -> totals if the direct .sum is absent
    def tot(op):
        k = f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum"
        if k in hdr:
            return float(r[hdr.index(k)].replace(",", ""))
        k2 = f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed"
        cyc = float(r[hdr.index("smsp__cycles_elapsed.avg")].replace(",", "")) if "smsp__cycles_elapsed.avg" in hdr else float(r[hdr.index("sm__cycles_elapsed.avg")].replace(",", ""))
        return float(r[hdr.index(k2)].replace(",", "")) * cyc
    try:
        dfma, dadd, dmul = tot("dfma"), tot("dadd"), tot("dmul")
        d["fp64_flops_per_launch"] = 2 * dfma + dadd + dmul
        d["fp64_thread_ops"] = {"dfma": dfma, "dadd": dadd, "dmul": dmul}
    except Exception as e:
        d["fp64_flops_error"] = str(e)
    out.append(d)
json.dump(out, open(outbase + ".json", "w"), indent=1)
with open(outbase + ".txt", "w") as f:
    for d in out:
        f.write("=" * 100 + "\n")
        for k, v in d.items():
            if isinstance(v, dict) and "value" in v:
                f.write(f"{k:70s} {v['value']} {v['unit']}\n")
            else:
                f.write(f"{k:70s} {v}\n")
print(open(outbase + ".txt").read())
