"""Bucket an `ncu --page source --csv` dump by barrier-delimited SASS segments (phases of a kernel)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
iS = hdr.index('# Samples'); iI = hdr.index('Instructions Executed'); iSrc = hdr.index('Source')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') or 'Stall' in h and 'Sampling' not in h]
tot_s = sum(int(r[iS]) for r in data); tot_i = sum(int(r[iI]) for r in data)
print('total samples', tot_s, 'total warp-instr', tot_i)
seg = []; cur = {'s': 0, 'i': 0, 'ops': collections.Counter(), 'start': 0, 'top': []}
for n, r in enumerate(data):
    op = r[iSrc].split()
    opn = op[0] if not op[0].startswith('@') else op[1]
    cur['s'] += int(r[iS]); cur['i'] += int(r[iI]); cur['ops'][opn.split('.')[0]] += int(r[iI])
    cur['top'].append((int(r[iS]), r[iSrc].strip()[:60]))
    if opn.startswith('BAR') or opn.startswith('EXIT'):
        cur['end'] = n; seg.append(cur); cur = {'s': 0, 'i': 0, 'ops': collections.Counter(), 'start': n + 1, 'top': []}
for k, sg in enumerate(seg):
    if sg['s'] < tot_s * 0.01 and sg['i'] < tot_i * 0.01: continue
    print(k, 'lines', sg['start'], sg['end'], 'samples %.1f%%' % (100 * sg['s'] / tot_s), 'instr %.1f%%' % (100 * sg['i'] / tot_i), dict(sg['ops'].most_common(6)))
    for s_, src in sorted(sg['top'], reverse=True)[:4]:
        print('      %5.2f%%  %s' % (100 * s_ / tot_s, src))
