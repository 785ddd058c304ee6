#!/usr/bin/env python
"""Dynamic opcode mix and stall breakdown per kernel from an `ncu --page source --csv` export.
usage: ncu_source_mix.py <source.csv> [top_lines]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1]).read().splitlines()))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 0
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
STALLS = ('stall_barrier', 'stall_long_sb', 'stall_math', 'stall_mio', 'stall_short_sb', 'stall_wait', 'stall_not_selected',
          'stall_selected', 'stall_dispatch', 'stall_lg', 'stall_no_inst', 'stall_branch_resolving')
for si, s in enumerate(secs):
    e = secs[si + 1] if si + 1 < len(secs) else len(rows)
    hdr = rows[s + 1]; idx = {h: i for i, h in enumerate(hdr)}
    ops = collections.Counter(); stalls = collections.Counter(); tot = 0; lines = []; by_op = collections.defaultdict(collections.Counter)
    for r in rows[s + 2:e]:
        if len(r) < len(hdr): continue
        src = r[idx['Source']].strip()
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+(\.\d+)?)', src)
        op = m.group(2) if m else src[:10]
        try: n = int(r[idx['Instructions Executed']])
        except ValueError: continue
        ops[op] += n; tot += n
        smp = 0
        for k in STALLS:
            try: v = int(r[idx[k]])
            except (ValueError, KeyError): v = 0
            stalls[k] += v; smp += v; by_op[op][k] += v
        lines.append((smp, n, r[idx['Address']], src))
    print(rows[s][1][:90], 'total warp instructions', tot)
    for op, n in ops.most_common(22): print('   %-12s %11d %5.1f%%' % (op, n, 100 * n / max(tot, 1)))
    ts = sum(stalls.values()) or 1
    print('   stall samples:', {k: round(100 * v / ts, 1) for k, v in stalls.most_common(8)})
    print('   stall samples by opcode (share of all samples; top reasons within the opcode):')
    for op, c in sorted(by_op.items(), key=lambda kv: -sum(kv[1].values()))[:10]:
        t = sum(c.values())
        print('     %-10s %5.1f%%  %s' % (op, 100 * t / ts, {k: round(100 * v / max(t, 1)) for k, v in c.most_common(4)}))
    if top:
        print('   hottest lines by stall samples:')
        for smp, n, a, src in sorted(lines, reverse=True)[:top]: print('     %6d %9d %s %s' % (smp, n, a, src[:100]))
