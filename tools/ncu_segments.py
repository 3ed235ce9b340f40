"""Developer tool: per-barrier-segment instruction / shared-memory-wavefront breakdown of one kernel from
`ncu -i X.ncu-rep --page source --csv` output (argv[1])."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
hdr = rows[hi]
data = rows[hi + 1:]
c = hdr.index
ie, src, wf, wfi, smp = c('Instructions Executed'), c('Source'), c('L1 Wavefronts Shared'), c('L1 Wavefronts Shared Ideal'), c('# Samples')
segs = []
cur = dict(n=0, wf=0, wfi=0, s=0, ops=collections.Counter(), lines=[])
for idx, r in enumerate(data):
    try:
        n = int(r[ie])
    except Exception:
        continue
    w, wi, s = int(r[wf] or 0), int(r[wfi] or 0), int(r[smp] or 0)
    ins = r[src]
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', ins)
    op = m.group(2) if m else ins
    cur['n'] += n; cur['wf'] += w; cur['wfi'] += wi; cur['s'] += s; cur['ops'][op.split('.')[0]] += n
    if w > 0:
        cur['lines'].append((w, wi, n, ins.strip()[:60]))
    if 'BAR.SYNC' in ins or idx == len(data) - 1:
        segs.append(cur)
        cur = dict(n=0, wf=0, wfi=0, s=0, ops=collections.Counter(), lines=[])
T = sum(s['n'] for s in segs); W = sum(s['wf'] for s in segs); S = sum(s['s'] for s in segs)
print('total instr', T, 'wf', W, 'samples', S)
for i, s in enumerate(segs):
    if s['n'] < 1000:
        continue
    print(f"seg{i}: instr {s['n']/1e6:.2f}M ({100*s['n']/T:.1f}%) wf {s['wf']/1e6:.2f}M ideal {s['wfi']/1e6:.2f}M samples {100*s['s']/S:.1f}%")
    print('    ', ', '.join(f'{k}:{v/1e6:.2f}' for k, v in s['ops'].most_common(14)))
    agg = collections.Counter(); aggi = collections.Counter(); cnt = collections.Counter()
    for w, wi, n, ins in s['lines']:
        key = re.sub(r'\[.*\]', '[]', ins)
        key = re.sub(r'R\d+', 'R', key)
        agg[key] += w; aggi[key] += wi; cnt[key] += n
    for k, v in agg.most_common(8):
        print(f'       {k:45s} wf {v/1e6:.2f}M ideal {aggi[k]/1e6:.2f}M instr {cnt[k]/1e6:.2f}M')
