"""tcgen05 / TMEM / TMA instruction census per kernel of libnca_b200.so (cuobjdump -sass), written to profiles/sass_census_r2.txt.
usage: python tools/sass_census.py > profiles/sass_census_r2.txt"""
import collections, os, re, subprocess
lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "video-stylization-with-nca_b200", "libnca_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
cnt, cur, i = collections.OrderedDict(), None, 0
for line in out.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = names[i].split("(")[0].replace("void ", ""); i += 1
        cnt[cur] = collections.Counter()
        continue
    if cur is None: continue
    for op, key in (("UTCHMMA", "UTCHMMA"), ("LDTM", "LDTM"), ("STTM", "STTM"), ("UTMALDG", "UTMALDG"), ("UBLKCP", "UBLKCP"), ("UTCBAR", "UTCBAR"),
                    ("LDL", "LDL/STL"), ("STL", "LDL/STL")):
        if re.search(r"\b" + op + r"\b|\b" + op + r"\.", line): cnt[cur][key] += 1
print("# tcgen05 / TMEM / TMA instructions per kernel in video-stylization-with-nca_b200/libnca_b200.so (round 2, final)")
print("# cuobjdump -sass libnca_b200.so, counted per function: UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG (TMA tensor load),")
print("# UBLKCP (bulk copy), UTCBAR (tcgen05.commit), LDL/STL (local memory: register spills).  Template arguments: dynca_fwd_tc2_kernel<NS, C, fc, ops_only>,")
print("# dynca_bwd_tc2_kernel<NS, C, fc>, dynca_*_bf16_kernel<NS, split precision (f16x3)>, enc_*_tc_kernel<C>; 0 = generic.")
for k, c in cnt.items():
    if c["UTCHMMA"] == 0: continue
    print("%-52s UTCHMMA=%d LDTM=%d STTM=%d UTMALDG=%d UBLKCP=%d UTCBAR=%d LDL/STL=%d" % (k, c["UTCHMMA"], c["LDTM"], c["STTM"], c["UTMALDG"], c["UBLKCP"], c["UTCBAR"], c["LDL/STL"]))
