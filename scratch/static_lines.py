"""Static code size per CUDA source line of one kernel (SASS instructions per line, from nvdisasm's line info): where a
kernel's code bloat is.  usage: python scratch/static_lines.py <cubin> <kernel-substring> [top_n]
(cubins: cuobjdump -xelf all dns_slam_b200/libdns_slam_b200.so)"""
import re, subprocess, sys, collections
cubin, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
cnt = collections.Counter(); cur = None; inside = False; total = 0
for line in out.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", line) or re.match(r"//-+ \.text\.(\S+)", line)
    if m:
        inside = kern in m.group(1); continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line) and cur:
        cnt[cur] += 1; total += 1
print(f"{total} instructions = {total * 16 / 1024:.1f} KB")
byfile = collections.Counter()
for (f, l), v in cnt.items(): byfile[f] += v
print({f: v for f, v in byfile.most_common(8)})
for (f, l), v in cnt.most_common(top):
    print(f"{v:6d} ({100 * v / total:4.1f} %)  {f}:{l}")
