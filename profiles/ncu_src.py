"""Summarises `ncu -i X.ncu-rep --page source --csv`: per-SASS-instruction samples, executed
count and average active threads; prints the hottest instructions."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot_s = sum(float(r[ix["# Samples"]] or 0) for r in body)
tot_i = sum(float(r[ix["Instructions Executed"]] or 0) for r in body)
tot_t = sum(float(r[ix["Thread Instructions Executed"]] or 0) for r in body)
print(f"instructions: {len(body)}  warp-inst executed: {tot_i:.3e}  thread-inst: {tot_t:.3e}  avg threads/inst: {tot_t/max(tot_i,1):.2f}  samples: {tot_s:.0f}")
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(f"{'#':>4} {'samples%':>8} {'exec%':>7} {'thr':>5}  stall(top)          sass")
for n, r in enumerate(body):
    r.append(n)
for r in sorted(body, key=lambda r: -float(r[ix["# Samples"]] or 0))[:top]:
    st = {k: float(r[ix[k]] or 0) for k in hdr if k.startswith("stall_") and "(" not in k}
    top2 = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{r[-1]:>4} {100*float(r[ix['# Samples']] or 0)/tot_s:8.2f} {100*float(r[ix['Instructions Executed']] or 0)/tot_i:7.2f} "
          f"{float(r[ix['Avg. Threads Executed']] or 0):5.1f}  {top2[0][0][6:]:>9}/{top2[1][0][6:]:<9} {r[ix['Source']][:80]}")
