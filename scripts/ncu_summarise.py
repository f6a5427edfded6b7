"""Turns the ncu artefacts of a gpurun call into the tracked summaries under profiles/:
    python scripts/ncu_summarise.py launches gpurun_out/X_launches.csv profiles/X_launches_ml10m   -> _table.md, _head.csv
    python scripts/ncu_summarise.py raw gpurun_out/X_prof.ncu-rep profiles/X_prof_raw.csv          -> `ncu --page raw --csv`
"""
import csv, re, subprocess, sys
from collections import OrderedDict

mode, src, dst = sys.argv[1:4]
if mode == "launches":
    lines = [l for l in open(src) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr, rows = rows[0], rows[1:]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*$", "", r[ik])[:160]
        v = float(r[iv].replace(",", ""))
        ms = v / 1e6 if r[iu] in ("ns", "nsecond") else (v / 1e3 if r[iu] in ("us", "usecond") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += ms
    tot = sum(a[1] for a in agg.values())
    with open(dst + "_table.md", "w") as f:
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.2f | %.1f%% | %.1f |\n" % (name, n, ms, 100 * ms / tot, 1e3 * ms / n))
        f.write("\n%d launches, %.1f ms in total\n" % (len(rows), tot))
    with open(dst + "_head.csv", "w") as f:
        f.writelines(lines[:120])
    print("%d launches, %.1f ms" % (len(rows), tot))
else:
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    open(dst, "w").write(out)
    print(len(out), "bytes")
