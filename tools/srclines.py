"""Per-CUDA-source-line instruction / stall-sample totals from `ncu -i X --page source --csv --print-source cuda,sass`.
usage: srclines.py dump.csv [kernel-substring] [top-n]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2] if len(sys.argv) > 2 else ''
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 60
fpath = func = None; hdr = None
agg = collections.defaultdict(lambda: [0, 0, ''])   # (func, file, line) -> [instr, samples, text]
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fpath = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': func = r[1]; continue
    if r[0] == 'Line No': hdr = r; iI = hdr.index('Instructions Executed'); iS = hdr.index('# Samples'); continue
    if hdr is None or len(r) < len(hdr) - 1: continue
    if r[0] == '': continue       # SASS row
    try: ins = int(r[iI]); smp = int(r[iS])
    except ValueError: continue
    if ins == 0 and smp == 0: continue
    k = (func, fpath, int(r[0])); agg[k][0] += ins; agg[k][1] += smp; agg[k][2] = r[1].strip()[:110]
funcs = sorted({k[0] for k in agg})
for f in funcs:
    if want not in f: continue
    items = [(k, v) for k, v in agg.items() if k[0] == f]
    ti = sum(v[0] for _, v in items); ts = sum(v[1] for _, v in items)
    print('==', f[:90], 'warp-instr', ti, 'samples', ts)
    byfile = collections.Counter()
    for k, v in items: byfile[k[1]] += v[0]
    print('   by file:', {a: '%.1f%%' % (100 * b / ti) for a, b in byfile.most_common()})
    for k, v in sorted(items, key=lambda kv: -kv[1][0])[:topn]:
        print('  %5.2f%% i %5.2f%% s  %s:%d  %s' % (100 * v[0] / ti, 100 * v[1] / max(ts, 1), k[1], k[2], v[2]))
