"""Per-kernel share of ONE step from an `ncu --metrics gpu__time_duration.sum --csv` launch list of tools/prof_step.py.
usage: python tools/step_breakdown.py launches.csv n_steps_in_list"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
nsteps = int(sys.argv[2])
t, n = collections.Counter(), collections.Counter()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = r[ki].split("(")[0][:70]
    t[name] += v / 1000.0
    n[name] += 1
tot = sum(t.values())
print(f"{'us':>9} {'share':>6} {'launches':>8}  kernel   (per step, {nsteps} steps in the list)")
for k, v in t.most_common():
    print(f"{v / nsteps:9.1f} {v / tot * 100:5.1f}% {n[k] / nsteps:8.1f}  {k}")
print(f"{tot / nsteps:9.1f} 100.0% {sum(n.values()) / nsteps:8.1f}  total")
