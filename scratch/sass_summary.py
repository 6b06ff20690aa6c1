"""SASS evidence per kernel of the built library: tcgen05 MMA (UTCHMMA), TMEM loads (LDTM), bulk copies (UBLKCP*), vector
reductions (RED.*128) -- usage: python scratch/sass_summary.py > profiles/<tag>_sass_tcgen05.txt"""
import re, subprocess, collections
sass = subprocess.run(["cuobjdump", "-sass", "dns_slam_b200/libdns_slam_b200.so"], capture_output=True, text=True).stdout
cur, cnt = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); cnt[cur] = collections.Counter(); continue
    if cur is None: continue
    for key, pat in (("UTCHMMA", r"\bUTCHMMA"), ("LDTM", r"\bLDTM"), ("UBLKCP", r"\bUBLKCP"), ("UTCBAR", r"\bUTCBAR"),
                     ("REDG.F32x4", r"\bREDG\S*F32x4"), ("REDG.F32x2", r"\bREDG\S*F32x2"),
                     ("SYNCS", r"\bSYNCS")):
        if re.search(pat, line): cnt[cur][key] += 1
print("# cuobjdump -sass of dns_slam_b200/libdns_slam_b200.so (sm_100a): instruction counts per kernel that uses the tensor / TMA path")
for k, c in sorted(cnt.items()):
    if c["UTCHMMA"] or c["UBLKCP"] or c["LDTM"]:
        print(k, " ".join(f"{n}={c[n]}" for n in ("UTCHMMA", "LDTM", "UBLKCP", "UTCBAR", "SYNCS", "REDG.F32x4", "REDG.F32x2") if c[n]))
