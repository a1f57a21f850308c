"""Linear SASS listing of one kernel from `ncu --page source --csv --print-source cuda,sass` with per-instruction executed counts,
stall samples and source lines (usage: sass_phases.py dump.csv kernel-substring out.txt)"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want, outp = sys.argv[2], sys.argv[3]
hdr = func = fpath = cur = None
out = {}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fpath = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': func = r[1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if func is None or want not in func: continue
    if r[0] != '': cur = (fpath, int(r[0])); continue
    if r[2] in ('...', '-'): continue
    try: addr = int(r[2], 16)
    except ValueError: continue
    iI = hdr.index('Instructions Executed'); iS = hdr.index('# Samples')
    ins = int(r[iI]) if r[iI].isdigit() else 0
    smp = int(r[iS]) if r[iS].isdigit() else 0
    st = {}
    for j, h in enumerate(hdr):
        if h.startswith('stall_') and 'Not Issued' not in h and j < len(r) and r[j].isdigit() and int(r[j]) > 0: st[h[6:]] = int(r[j])
    if addr not in out or ins > 0: out[addr] = (r[3].strip(), ins, smp, cur, st)
base = min(out)
with open(outp, 'w') as f:
    for a in sorted(out):
        s, ins, smp, cur, st = out[a]
        f.write('%6x %8d %5d %-22s %-70s %s\n' % (a - base, ins, smp, '%s:%d' % (cur[0][:14], cur[1]), s[:70], st if st else ''))
print(len(out), 'instructions,', sum(v[1] for v in out.values()), 'executed,', sum(v[2] for v in out.values()), 'samples')
