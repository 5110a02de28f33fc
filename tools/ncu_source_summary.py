"""Summarises `ncu -i X.ncu-rep --page source --csv --print-source sass` for the first kernel: stall mix, instruction mix, hottest SASS lines."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hdr = None; data = []; nk = 0; name = ""
for r in rows:
    if r and r[0] == "Kernel Name":
        nk += 1
        if nk == which: name = r[1]
        if nk > which: break
        continue
    if nk != which: continue
    if r and r[0] == "Address": hdr = r; continue
    if hdr and len(r) >= len(hdr) - 2: data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("kernel", name, "sass rows", len(data), "samples", tot)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stalls}
for s, v in sorted(agg.items(), key=lambda x: -x[1])[:10]: print(f"  {s:28s} {v:8d} {100*v/tot:5.1f}%")
mix = collections.Counter(); smp = collections.Counter()
for r in data:
    op = [o for o in r[ix["Source"]].split() if not o.startswith('@')][0].split('.')[0]
    mix[op] += int(r[ix["Instructions Executed"]]); smp[op] += int(r[ix["# Samples"]])
t = sum(mix.values())
print("instruction mix (warp instructions executed, share; stall samples share)")
for o, v in mix.most_common(22): print(f"  {o:12s} {v:12d} {100*v/t:5.1f}%   samples {100*smp[o]/tot:5.1f}%")
print("hottest lines")
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:25]:
    top = sorted(((int(r[ix[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"  {int(r[ix['# Samples']]):6d} {r[ix['Source']].strip()[:70]:70s} {top}")
