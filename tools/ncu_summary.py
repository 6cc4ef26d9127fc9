"""Summarise an .ncu-rep (one kernel) into a small JSON + text block for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name"""
import csv, collections, io, json, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
summ = {}
for h, u, v in zip(hdr, units, vals):
    key = h.split(".TriageCompute.")[-1] if "TriageCompute" in h else h
    if key in want:
        summ[key] = f"{v} {u}".strip()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot_i = sum(f(r, "Instructions Executed") for r in data)
tot_s = sum(f(r, "# Samples") for r in data)
ops = collections.Counter()
for r in data:
    t = r[ix["Source"]].split()
    op = (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?"))
    ops[op] += f(r, "Instructions Executed")
st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
stalls = {h[6:]: sum(f(r, h) for r in data) for h in st}
ssum = sum(stalls.values()) or 1
summ["opcode_mix_pct"] = {k: round(100 * v / tot_i, 2) for k, v in ops.most_common(12)}
summ["stall_reasons_pct"] = {k: round(100 * v / ssum, 1) for k, v in sorted(stalls.items(), key=lambda x: -x[1])[:8]}
summ["samples"] = tot_s
json.dump(summ, open(out + ".json", "w"), indent=1)
print(json.dumps(summ, indent=1))
