"""Aggregate an ncu report's source page per CUDA source line: instructions executed and stall samples.
usage: python scratch/src_lines.py <rep> <kernel-regex> [top_n]"""
import csv, subprocess, sys, io
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None
agg = {}
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 10 or r[0] == "":
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    i_inst = hdr.index("Instructions Executed")
    i_samp = hdr.index("# Samples")
    key = (cur_file, line)
    a = agg.setdefault(key, [0, 0, r[1]])
    def num(v):
        try: return int(v)
        except ValueError: return 0
    a[0] += num(r[i_inst])
    a[1] += num(r[i_samp])
tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[1] for a in agg.values())
print(f"total warp instr {tot_i}, samples {tot_s}")
print("--- by samples")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*a[1]/max(tot_s,1):5.1f}% smp {100*a[0]/max(tot_i,1):5.1f}% ins  {f}:{l}  {a[2].strip()[:110]}")

print("--- by instructions")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/max(tot_i,1):5.1f}% ins {100*a[1]/max(tot_s,1):5.1f}% smp  {f}:{l}  {a[2].strip()[:110]}")
