"""Summarise an .ncu-rep (one kernel launch) into JSON + the hottest source lines:  python scripts/ncu_summary.py rep out_prefix"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, v = rows[0], rows[-1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_wait",
        "smsp__pcsamp_warps_issue_stalled_short_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_mio_throttle",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_sleeping"]
summ = {}
for i, n in enumerate(h):
    if n in want:
        summ[n] = {"unit": rows[1][i], "value": v[i]}
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur, lines = None, []
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1]; continue
    if len(r) > 8 and r[0].isdigit():
        try:
            lines.append((int(r[6]), int(r[7]), cur.split("/")[-1], int(r[0]), r[1].strip()[:140]))
        except ValueError:
            pass
tot = sum(l[0] for l in lines) or 1
json.dump(summ, open(out + "_ncu_summary.json", "w"), indent=1)
with open(out + "_hot_lines.txt", "w") as f:
    f.write(f"# warp-stall samples per source line (ncu --set full --import-source on), total {tot}\n")
    for l in sorted(lines, reverse=True)[:30]:
        f.write(f"{l[0]:7d} {100 * l[0] / tot:5.1f}%  inst {l[1]:10d}  {l[2]}:{l[3]}  {l[4]}\n")
print(json.dumps({k: v["value"] for k, v in summ.items()}, indent=0)[:1500])
