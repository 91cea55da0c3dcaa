"""Aggregate an `ncu --page source --csv --print-source sass` dump by opcode: stall samples, instructions, smem wavefronts."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix['# Samples']]) for r in data)
print('total samples', tot, 'instr', len(data))
stall = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {s: sum(int(r[ix[s]] or 0) for r in data) for s in stall}
print(sorted(agg.items(), key=lambda x: -x[1])[:8])
op = collections.Counter(); opi = collections.Counter(); wf = collections.Counter(); wfi = collections.Counter()
for r in data:
    o = [t for t in r[ix['Source']].split() if not t.startswith('@')][0]
    op[o] += int(r[ix['# Samples']]); opi[o] += int(r[ix['Instructions Executed']])
    wf[o] += int(r[ix['L1 Wavefronts Shared']] or 0); wfi[o] += int(r[ix['L1 Wavefronts Shared Ideal']] or 0)
for o, c in op.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print(f"{o:28s} samples {c:7d} {100*c/tot:5.1f}%  inst {opi[o]:12d} smem wf {wf[o]:11d} ideal {wfi[o]:11d}")
print('total inst', sum(opi.values()))
