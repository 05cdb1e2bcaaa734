"""Device-side timeline of the CG loop from the STAN_CG_TRACE dumps (one csv per rank: ns, iteration, event code).

    STAN_CG_TRACE=8192 STAN_CG_TRACE_FROM=200 STAN_CG_TRACE_DIR=gpurun_out/trace torchrun ... tools/multi_iter_time.py ...
    python tools/cg_timeline.py gpurun_out/trace > profiles/r02_cg_timeline_8gpu.md

Every kernel of the loop stamps %globaltimer from one thread at its begin and at the points where the local part
of a reduction is done and where the cross-rank sum has arrived.  The table lists, per transition between
consecutive events of one rank, the median and mean duration over the traced iterations — i.e. where the time of
an iteration goes: kernel bodies, waiting for the other ranks, and gaps between kernels."""
import collections
import glob
import os
import statistics
import sys

NAMES = {1: "spmv begin", 2: "spmv local sum done", 3: "spmv end (p.Ap from all ranks)", 4: "update begin", 5: "update local sums done",
         6: "update end (r.r, r.z from all ranks)", 7: "direction begin", 8: "halo push begin", 9: "halo flags raised",
         10: "halo wait begin", 11: "halo wait end", 12: "refresh begin", 13: "refresh end"}
d = sys.argv[1] if len(sys.argv) > 1 else "."
files = sorted(glob.glob(os.path.join(d, "stan_cg_trace_rank*.csv")))
per_rank = {}
for f in files:
    rank = int(f.rsplit("rank", 1)[1].split(".")[0])
    ev = sorted(tuple(int(v) for v in line.split(",")) for line in open(f) if line.strip())
    per_rank[rank] = ev
print(f"# CG timeline, {len(files)} rank(s), from `{d}` (tools/cg_timeline.py)\n")
agg = collections.defaultdict(list)
iters = {}
for rank, ev in per_rank.items():
    for (t0, k0, c0), (t1, k1, c1) in zip(ev, ev[1:]):
        agg[(c0, c1)].append((t1 - t0) / 1e3)
    marks = [t for t, k, c in ev if c == 7]
    if len(marks) > 2:
        iters[rank] = (marks[-1] - marks[0]) / 1e3 / (len(marks) - 1)
print("| from | to | count | median us | mean us |")
print("|---|---|---|---|---|")
order = sorted(agg, key=lambda k: -statistics.mean(agg[k]) * len(agg[k]))
total = sum(sum(v) for v in agg.values())
for k in order:
    v = agg[k]
    if len(v) < 8:
        continue
    print(f"| {NAMES.get(k[0], k[0])} | {NAMES.get(k[1], k[1])} | {len(v)} | {statistics.median(v):.1f} | {statistics.mean(v):.1f} |")
if iters:
    print(f"\nIteration period (direction begin to direction begin), mean over ranks: {statistics.mean(iters.values()):.1f} us "
          f"(min {min(iters.values()):.1f}, max {max(iters.values()):.1f}).")
