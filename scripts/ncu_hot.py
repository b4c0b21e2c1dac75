"""usage: ncu_hot.py <source-page.csv> [top]  -> totals of stall reasons, instruction regions and the hottest SASS lines."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot_s = sum(f(r, "# Samples") for r in data); tot_i = sum(f(r, "Instructions Executed") for r in data)
tot_t = sum(f(r, "Thread Instructions Executed") for r in data)
print(f"samples {tot_s:.0f}, warp instr {tot_i:.3e}, thread instr {tot_t:.3e}, avg lanes {tot_t/tot_i:.2f}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(f(r, s) for r in data) for s in stalls}
print("stall samples:", ", ".join(f"{k[6:]} {100*v/tot_s:.1f}%" for k, v in sorted(agg.items(), key=lambda x: -x[1]) if v > 0.005 * tot_s))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print("\nhottest instructions (by samples):")
for n, r in sorted(enumerate(data), key=lambda x: -f(x[1], "# Samples"))[:top]:
    main = max(stalls, key=lambda s: f(r, s))
    print(f"{n:4d} {100*f(r,'# Samples')/tot_s:5.2f}%  exec {f(r,'Instructions Executed'):.2e} lanes {f(r,'Avg. Threads Executed'):4.1f}  {main[6:]:12s} {r[ix['Source']].strip()[:70]}")
# cumulative by position, to see regions
print("\ncumulative by position (every 40 instructions): idx, % samples, % warp instr")
cs = ci = 0
for n, r in enumerate(data):
    cs += f(r, "# Samples"); ci += f(r, "Instructions Executed")
    if n % 40 == 39 or n == len(data) - 1:
        print(f"  ..{n:4d}: samples {100*cs/tot_s:5.1f}%  instr {100*ci/tot_i:5.1f}%")
