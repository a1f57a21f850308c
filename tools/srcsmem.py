"""Shared-memory wavefronts (total / excessive) per CUDA source line from `ncu --page source --csv --print-source cuda,sass`."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ''
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
fpath = func = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, ''])
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fpath = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': func = r[1]; continue
    if r[0] == 'Line No': hdr = r; iW = hdr.index('L1 Wavefronts Shared'); iE = hdr.index('L1 Wavefronts Shared Excessive'); continue
    if hdr is None or len(r) < len(hdr) - 1 or r[0] == '': continue
    try: w = int(r[iW]); e = int(r[iE])
    except ValueError: continue
    if w == 0: continue
    k = (func, fpath, int(r[0])); agg[k][0] += w; agg[k][1] += e; agg[k][2] = r[1].strip()[:100]
for f in sorted({k[0] for k in agg}):
    if want not in f: continue
    items = [(k, v) for k, v in agg.items() if k[0] == f]
    tw = sum(v[0] for _, v in items); te = sum(v[1] for _, v in items)
    print('==', f[:80], 'smem wavefronts', tw, 'excessive', te)
    for k, v in sorted(items, key=lambda kv: -kv[1][0])[:topn]:
        print('  %5.2f%% w (%5.2f%% exc)  %s:%d  %s' % (100 * v[0] / tw, 100 * v[1] / max(v[0], 1), k[1], k[2], v[2]))
