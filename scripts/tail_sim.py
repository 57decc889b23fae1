"""Development aid (profiles/r02_tail.md): list-scheduling simulation of one bench launch on the measured iteration counts
(gpurun_out/iters_by_nsqp.json from scripts/dump_iters.py): how much of the kernel time is the tail of the last wave, what a
longest-first order would give, and whether splitting the launch after k SQP iterations to re-order by the observed counts pays."""
import heapq, json, os, sys
import numpy as np
from scipy.stats import spearmanr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = json.load(open(os.path.join(ROOT, "gpurun_out", "iters_by_nsqp.json")))
Q = {int(n): np.array(v)[:, 0].astype(float) for n, v in d.items()}
S = {int(n): np.array(v)[:, 1].astype(float) for n, v in d.items()}
CQ, CS, SLOTS, GHZ = 51.8e3, 29.4e3, 592, 1.965e6  # cycles per IPM / SQP iteration under load (profiles/r02_phases_cfg2_loaded.txt), 148 SMs x 4


def sched(dur, order):
    h = [0.0] * SLOTS
    heapq.heapify(h)
    end = 0.0
    for i in order:
        t = heapq.heappop(h) + dur[i]
        end = max(end, t)
        heapq.heappush(h, t)
    return end / GHZ


tot = Q[10]
full = tot * CQ + S[10] * CS
print("| schedule | makespan, ms |\n|---|---|")
print(f"| sum of durations / {SLOTS} slots (no tail) | {full.sum() / SLOTS / GHZ:.2f} |")
print(f"| launch order (what the hardware does) | {sched(full, range(len(full))):.2f} |")
print(f"| longest first (needs the durations in advance) | {sched(full, np.argsort(-full)):.2f} |")
print(f"| shortest first | {sched(full, np.argsort(full)):.2f} |")
print("\n| split after k SQP iterations | Spearman(IPM iterations so far, remaining) | first kernel + rest ordered by the observed counts, ms | rest ordered by its true length, ms |\n|---|---|---|---|")
for n in (1, 2, 3, 4):
    A = Q[n] * CQ + S[n] * CS
    B = (tot - Q[n]) * CQ + (S[10] - S[n]) * CS
    tA = sched(A, range(len(A)))
    print(f"| {n} | {spearmanr(Q[n], tot - Q[n]).correlation:.2f} | {tA + sched(B, np.argsort(-Q[n])):.2f} | {tA + sched(B, np.argsort(-B)):.2f} |")
print(f"\nlongest solve {full.max() / GHZ:.2f} ms, mean {full.mean() / GHZ:.2f} ms; IPM iterations per solve: percentiles 0/10/50/90/99/100 = {np.percentile(tot, [0, 10, 50, 90, 99, 100])}")
