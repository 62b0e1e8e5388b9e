"""Summarise an ncu report (run here, no GPU needed): key raw metrics + per-source-line instruction shares.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_rollout_full.txt
"""
import collections
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread
launch__shared_mem_per_block_dynamic launch__occupancy_limit_shared_mem launch__occupancy_limit_registers launch__waves_per_multiprocessor
sm__warps_active.avg.pct_of_peak_sustained_active sm__cycles_elapsed.avg smsp__cycles_active.avg smsp__inst_executed.sum
sm__inst_executed.avg.per_cycle_elapsed smsp__issue_active.avg.pct_of_peak_sustained_active smsp__thread_inst_executed_per_inst_executed.ratio
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
smsp__sass_inst_executed_op_shared_ld.sum smsp__sass_inst_executed_op_shared_st.sum
dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
smsp__average_warp_latency_per_inst_issued.ratio smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__sass_thread_inst_executed_op_ffma_pred_on.sum smsp__sass_thread_inst_executed_op_fadd_pred_on.sum smsp__sass_thread_inst_executed_op_fmul_pred_on.sum""".split()
lines = []
for r in rows[2:]:
    lines.append(f"== launch: {r[hdr.index('Kernel Name')][:80]}  grid {r[hdr.index('Grid Size')] if 'Grid Size' in hdr else ''}")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            lines.append(f"{w:90s} {r[i]:>22s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur, h, agg, text = None, None, collections.defaultdict(lambda: [0, 0]), {}
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 5 and r[0] == "Line No":
        h = r
        ie, isamp = h.index("Instructions Executed"), h.index("# Samples")
    elif h and len(r) == len(h) and r[0].isdigit() and r[2] == "-":
        try:
            agg[(cur, int(r[0]))][0] += int(r[ie]); agg[(cur, int(r[0]))][1] += int(r[isamp]); text[(cur, int(r[0]))] = r[1]
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
lines.append("")
lines.append(f"== per-source-line share of executed warp instructions / stall samples (top 40 of {len(agg)} lines, total {tot})")
for (f, ln), v in sorted(agg.items(), key=lambda x: -x[1][0])[:40]:
    lines.append(f"{100*v[0]/tot:5.2f}% inst {100*v[1]/ts:5.2f}% smp  {f}:{ln}: {text[(f, ln)].strip()[:100]}")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:45]))
