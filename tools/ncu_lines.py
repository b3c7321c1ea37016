"""Correlate an ncu report's per-SASS-instruction counters with source lines (via nvdisasm -g).
usage: python tools/ncu_lines.py <report.ncu-rep> <object.o> <kernel-substring> [source.cu]"""
import collections, csv, io, re, subprocess, sys, os, tempfile
rep, obj, pat = sys.argv[1:4]
srcfile = sys.argv[4] if len(sys.argv) > 4 else None
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith("\t.section\t.text.") and pat in l)
insts, line = [], None
for l in sass[start + 1:]:
    if l.startswith("\t.section") and insts:
        break
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        insts.append((line, m.group(2)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]
ii, sm, ti = h.index("Instructions Executed"), h.index("# Samples"), h.index("Thread Instructions Executed")
data = [(int(r[ii]), int(r[sm]), int(r[ti])) for r in rows[2:] if len(r) > ii and r[ii].isdigit()]
assert len(data) == len(insts), (len(data), len(insts))
by, bs, bt = collections.Counter(), collections.Counter(), collections.Counter()
for (ln, op), (c, s, t) in zip(insts, data):
    by[ln] += c; bs[ln] += s; bt[ln] += t
tot, ts = sum(by.values()), sum(bs.values())
src = open(srcfile).read().split("\n") if srcfile else None
print("SASS insts", len(insts), "executed warp-insts", tot, "samples", ts, "avg thr/inst %.1f" % (sum(bt.values()) / tot))
for ln, c in by.most_common(int(os.environ.get("TOP", "40"))):
    txt = ""
    if src and ln and ln[0] == os.path.basename(srcfile):
        txt = src[ln[1] - 1].strip()[:90]
    print(f"{c/tot*100:5.1f}% inst {bs[ln]/max(ts,1)*100:5.1f}% samp thr/inst {bt[ln]/max(c,1):4.1f}  {ln}: {txt}")
bands = collections.Counter()
for ln, c in by.items():
    t = bt[ln] / max(c, 1)
    bands["conv(>=24)" if t >= 24 else "partial(8-24)" if t >= 8 else "divergent(<8)"] += c
print({k: f"{v/tot*100:.1f}%" for k, v in bands.items()})
