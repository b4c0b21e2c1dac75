"""usage: ncu_lines.py <report.ncu-rep> <kernel-regex> [top]  -> hottest CUDA source lines (samples, warp instructions, lanes)
from `ncu --page source --print-source cuda,sass`."""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
lines = {}; fname = ""; hdr = None; cur = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None: continue
    def f(k):
        try: return float(r[hdr[k]])
        except Exception: return 0.0
    if r[0].strip():                      # a CUDA source line: its own totals
        key = (fname, int(r[0])); lines[key] = dict(src=r[1].strip(), s=f("# Samples"), i=f("Instructions Executed"), t=f("Thread Instructions Executed"))
S = sum(v["s"] for v in lines.values()); I = sum(v["i"] for v in lines.values())
print(f"samples {S:.0f}, warp instructions {I:.3e}")
for (fn, ln), v in sorted(lines.items(), key=lambda x: -x[1]["i"])[:top]:
    if v["i"] == 0: continue
    print(f"{fn}:{ln:5d} instr {100*v['i']/I:5.1f}% samples {100*v['s']/S:5.1f}% lanes {v['t']/max(v['i'],1):4.1f}  {v['src'][:100]}")
