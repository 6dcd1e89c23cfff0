"""Summarise `ncu --page source --csv` output: per-kernel opcode mix and hottest SASS lines."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
kern = None; hdr = None; blocks = {}
for r in rows:
    if r and r[0] == "Kernel Name": kern = r[1][:60]; blocks[kern] = []; hdr = None; continue
    if r and r[0] == "Address": hdr = r; continue
    if hdr and len(r) == len(hdr): blocks[kern].append(r)
for kern, data in blocks.items():
    isrc = hdr.index('Source'); iex = hdr.index('Instructions Executed'); ismp = hdr.index('# Samples')
    tot = sum(int(r[iex]) for r in data); ts = sum(int(r[ismp]) for r in data)
    print('=====', kern, 'warp-inst', tot, 'sass lines', len(data), 'samples', ts)
    byop = collections.Counter(); smp = collections.Counter()
    for r in data:
        t = r[isrc].split(); op = t[1] if t[0].startswith('@') else t[0]; op = op.rstrip(';')
        op = '.'.join(op.split('.')[:2]) if op.startswith(('LD','ST','ATOM','RED','BAR','MATCH','VOTE','SHFL')) else op.split('.')[0]
        byop[op] += int(r[iex]); smp[op] += int(r[ismp])
    for op, c in byop.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 22):
        print(f"  {op:14s} {c:12d} {100*c/max(tot,1):5.1f}%  samples {smp[op]:7d} {100*smp[op]/max(ts,1):5.1f}%")
