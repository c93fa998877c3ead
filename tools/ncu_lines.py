"""Per-CUDA-source-line instruction / stall-sample shares from an ncu report.
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [min_pct]"""
import csv, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.7
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
lines = []
fname = ""
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        ii = hdr.index("Instructions Executed"); sm = hdr.index("# Samples"); ti = hdr.index("Thread Instructions Executed")
    elif hdr and r and r[0].isdigit() and len(r) > ii:
        try:
            lines.append((fname, int(r[0]), r[1].strip(), float(r[ii]), float(r[sm]), float(r[ti])))
        except ValueError:
            pass
tot = sum(l[3] for l in lines); tots = sum(l[4] for l in lines)
print(f"total warp-instr {tot:.0f}  samples {tots:.0f}")
for f, n, src, ins, smp, tins in lines:
    if ins > minpct / 100 * tot or smp > minpct / 100 * tots:
        print(f"{100*ins/tot:5.1f}% instr {100*smp/max(tots,1):5.1f}% stall  thr/instr {tins/max(ins,1):4.1f}  {f}:{n:<4} {src[:95]}")
