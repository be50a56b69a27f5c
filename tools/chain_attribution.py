#!/usr/bin/env python
"""Where a step's wall time goes, from a tools/timeline.py CSV (one CUDA-graph replay, CUPTI kernel records).

CUPTI durations of kernels launched with programmatic dependent launch include the time they spend in
`griddepcontrol.wait` behind their predecessor, so summing durations double-counts. This walks each stream in start order and
charges every kernel only the time by which it ADVANCES the stream's frontier (end - max(previous frontier, start)); idle
gaps of the stream are reported separately. The forward pass is one chain (main stream); the backward pass is the
data-gradient / BN chain with the weight-gradient GEMMs on a side stream next to it.

    python tools/chain_attribution.py profiles/r2_timeline_step_graph_replay_final.csv"""
import collections
import csv
import sys


def main():
    rows = list(csv.DictReader(open(sys.argv[1])))
    for r in rows:
        r["s"], r["d"] = float(r["start_us"]), float(r["dur_us"])
        r["e"] = r["s"] + r["d"]
    t0 = next(r["s"] for r in rows if "weights_batched" in r["name"])      # first kernel of the step proper
    rows = [r for r in rows if r["s"] >= t0 - 20]
    t_bwd = min(r["s"] for r in rows if "softmax_nll_bwd" in r["name"])
    end = max(r["e"] for r in rows)
    print("step %.0f us: forward %.0f us, backward %.0f us" % (end - t0, t_bwd - t0, end - t_bwd))
    streams = collections.Counter(r["stream"] for r in rows)
    for st, n in streams.most_common():
        ks = sorted((r for r in rows if r["stream"] == st), key=lambda r: r["s"])
        if n < 20:
            continue
        fam = collections.OrderedDict()
        front, idle = ks[0]["s"], 0.0
        for k in ks:
            idle += max(0.0, k["s"] - front)
            adv = max(0.0, k["e"] - max(front, k["s"]))
            front = max(front, k["e"])
            name = k["name"].split("(")[0][:34]
            phase = "fwd" if k["s"] < t_bwd else "bwd"
            f = fam.setdefault((phase, name), [0, 0.0])
            f[0] += 1
            f[1] += adv
        print("\nstream %s: %d kernels, %.0f .. %.0f us, idle between its kernels %.0f us" % (st, n, ks[0]["s"] - t0, front - t0, idle))
        for (phase, name), (c, us) in sorted(fam.items(), key=lambda kv: -kv[1][1])[:14]:
            print("  %-4s %-36s %4d launches %8.1f us  (%.1f per launch)" % (phase, name, c, us, us / c))


if __name__ == "__main__":
    main()
