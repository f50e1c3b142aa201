#!/usr/bin/env python
"""ncu_source_top.py <source.csv[.gz]> [kernel-index] -- per-kernel hot spots of an exported `--page source --csv` file:
instructions with the most excessive shared-memory wavefronts and the most stall samples."""
import csv, gzip, sys
path = sys.argv[1]
want = int(sys.argv[2]) if len(sys.argv) > 2 else None
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
kernels = []
cur = None
for row in csv.reader(f):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        kernels.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = row
    elif cur is not None and row:
        cur["rows"].append(row)
for ki, k in enumerate(kernels):
    if want is not None and ki != want:
        continue
    h = k["hdr"]
    ix = {n: i for i, n in enumerate(h)}
    def num(r, n):
        try:
            return float(r[ix[n]])
        except Exception:
            return 0.0
    rows = k["rows"]
    tot_s = sum(num(r, "# Samples") for r in rows)
    tot_w = sum(num(r, "L1 Wavefronts Shared") for r in rows)
    tot_i = sum(num(r, "L1 Wavefronts Shared Ideal") for r in rows)
    print(f"=== [{ki}] {k['name'][:110]}\n    samples {tot_s:.0f}  smem wavefronts {tot_w:.3g} (ideal {tot_i:.3g})  instr {len(rows)}")
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    agg = {n: sum(num(r, n) for r in rows) for n in stalls}
    print("    stalls:", ", ".join(f"{n[6:]} {v / max(tot_s, 1) * 100:.1f}%" for n, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    print("  -- top excessive smem wavefronts")
    for r in sorted(rows, key=lambda r: -num(r, "L1 Wavefronts Shared Excessive"))[:12]:
        if num(r, "L1 Wavefronts Shared Excessive") <= 0:
            break
        print(f"    {r[ix['Source']][:70]:70s} wf {num(r, 'L1 Wavefronts Shared'):.3g} ideal {num(r, 'L1 Wavefronts Shared Ideal'):.3g} nway {r[ix['L1 Conflicts Shared N-Way']]}")
    print("  -- top stall samples")
    for r in sorted(rows, key=lambda r: -num(r, "# Samples"))[:14]:
        top = max(stalls, key=lambda n: num(r, n))
        print(f"    {r[ix['Source']][:70]:70s} {num(r, '# Samples') / max(tot_s, 1) * 100:5.1f}%  {top[6:]}")
