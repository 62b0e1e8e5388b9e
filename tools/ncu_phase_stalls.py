"""Where one rollout launch spends its warp time, from the SASS page of an `ncu --set full --import-source on` report
(run here, no GPU needed).  The kernel's CTA-wide phase barriers (BAR.SYNC) cut the SASS into stretches that follow the
phases of a sub-step; a warp waiting at a barrier is sampled on the first instruction AFTER it, so each stretch's
"barrier" samples are the wait for the slowest warp of the stretch before it.

    python tools/ncu_phase_stalls.py gpurun_out/r2zz_prof.ncu-rep profiles/r2zz_final_phase_stalls.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
rows = [r for r in rows[rows.index(hdr) + 1:] if len(r) == len(hdr)]
C = {h: i for i, h in enumerate(hdr)}
STALLS = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]


def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


def opcode(ins):
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", ins.strip())
    return m.group(2) if m else "?"


total = sum(num(r[C["# Samples"]]) for r in rows)
total_ex = sum(num(r[C["Instructions Executed"]]) for r in rows)
# stretches: cut after every BAR.SYNC / EXIT (the code after the kernel's EXIT are its __noinline__ functions)
segs, cur = [], None
for k, r in enumerate(rows):
    if cur is None:
        cur = dict(first=k, n=0, samp=0, ex=0, thr=0, st=collections.Counter(), ops=collections.Counter(), wait=0, shfl=0, calls=0)
        # the wait at the barrier this stretch starts behind: charged to the first instruction that has to issue after it
        # (within its first four instructions; a stretch of one instruction is the barrier loop of a warp without a tile)
    ins = r[C["Source"]]
    ex = num(r[C["Instructions Executed"]])
    cur["n"] += 1
    cur["samp"] += num(r[C["# Samples"]])
    cur["ex"] += ex
    cur["thr"] += num(r[C["Thread Instructions Executed"]])
    cur["ops"][opcode(ins)] += ex
    cur["shfl"] += "SHFL" in ins
    cur["calls"] += "CALL" in ins
    for s in STALLS:
        cur["st"][s] += num(r[C[s]])
    if "BAR.SYNC" in ins or re.search(r"\bEXIT\b", ins):
        cur["wait"] = sum(num(q[C["stall_barrier"]]) for q in rows[cur["first"]:min(cur["first"] + 4, k + 1)])
        segs.append(cur)
        cur = None
if cur:
    cur["wait"] = sum(num(q[C["stall_barrier"]]) for q in rows[cur["first"]:cur["first"] + 4])
    segs.append(cur)

L = [f"# Per-stretch warp-time profile of one rollout launch ({rep.split('/')[-1]})", "",
     "Stretches are the SASS between two CTA phase barriers, in address order; `share` = warp samples in the stretch / all",
     "samples (ncu's PC sampling, so it is warp time, waiting included); `barrier wait` = barrier-stall samples on the first",
     "instructions after the barrier the stretch starts behind = the time its warps stood waiting for the slowest warp of the previous",
     "phase; `lanes` = threads per executed warp instruction.", "",
     f"Total: {total} samples, {total_ex / 1e9:.2f} G warp instructions.", "",
     "| stretch | static instr | executed (M) | share of warp time | of which barrier wait | lanes | SHFL / CALL sites | top stall reasons (share of the stretch) | top opcodes |",
     "|---|---|---|---|---|---|---|---|---|"]
for i, s in enumerate(segs):
    if s["samp"] < 0.003 * total:
        continue
    top = sorted(s["st"].items(), key=lambda kv: -kv[1])[:4]
    ops = s["ops"].most_common(5)
    L.append(f"| {i} (instr {s['first']}..) | {s['n']} | {s['ex'] / 1e6:.0f} | {100 * s['samp'] / total:.1f} % | {100 * s['wait'] / total:.1f} % | "
             f"{s['thr'] / max(1, s['ex']):.1f} | {s['shfl']} / {s['calls']} | "
             + ", ".join(f"{k[6:]} {100 * v / max(1, s['samp']):.0f} %" for k, v in top) + " | "
             + ", ".join(f"{k} {100 * v / max(1, s['ex']):.0f} %" for k, v in ops) + " |")
allst = collections.Counter()
for s in segs:
    allst.update(s["st"])
L += ["", "Whole kernel, stall reasons as share of all samples: "
      + ", ".join(f"{k[6:]} {100 * v / total:.1f} %" for k, v in sorted(allst.items(), key=lambda kv: -kv[1])[:8]) + "."]
allops = collections.Counter()
for s in segs:
    allops.update(s["ops"])
L += ["", "Whole kernel, executed warp instructions by opcode: "
      + ", ".join(f"{k} {100 * v / total_ex:.1f} %" for k, v in allops.most_common(16)) + "."]
open(out, "w").write("\n".join(L) + "\n")
print("\n".join(L))
