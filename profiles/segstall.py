"""Per-segment stall-reason breakdown of an `ncu --page source --csv` dump (segments = barrier-delimited SASS ranges)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
iI = hdr.index('Instructions Executed'); iSrc = hdr.index('Source'); iS = hdr.index('# Samples')
cols = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iS]) for r in data)
seg = collections.Counter(); segi = 0; acc = collections.defaultdict(collections.Counter); ins = collections.Counter()
for r in data:
    for i, h in cols:
        acc[segi][h] += int(r[i] or 0)
    ins[segi] += int(r[iI])
    op = r[iSrc].split(); opn = op[1] if op[0].startswith('@') else op[0]
    if opn.startswith('BAR') or opn.startswith('EXIT'): segi += 1
for k in sorted(acc):
    t = sum(acc[k].values())
    if t < 0.01 * tot: continue
    print(k, 'samples %.1f%% instr %d' % (100 * t / tot, ins[k]), {h[6:]: '%.0f%%' % (100 * v / t) for h, v in acc[k].most_common(6)})
