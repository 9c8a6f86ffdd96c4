"""Summarise an ncu source-page CSV (`ncu -i X.ncu-rep --page source --csv`) of a tc kNN kernel by code region.

Regions are delimited by landmark SASS instructions: first LDTM (stream start), the exact re-rank's
first LDG.E.128 run after the last VOTE, the consumer after the certificate's atomics.
"""
import csv, sys, collections

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; data = rows[2:]
si = hdr.index('# Samples'); ii = hdr.index('Instructions Executed'); src = hdr.index('Source')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]

def op(r):
    t = r[src].split()
    o = t[1] if t[0].startswith('@') else t[0]
    return o

idx_ldtm = [k for k, r in enumerate(data) if op(r).startswith('LDTM')]
idx_vote = [k for k, r in enumerate(data) if op(r).startswith('VOTE.ANY')]
idx_ldg = [k for k, r in enumerate(data) if op(r).startswith('LDG.E.ENL2.256') or op(r).startswith('LDG.E.128')]
idx_atom = [k for k, r in enumerate(data) if op(r).startswith('ATOM') or op(r).startswith('RED')]
s0 = idx_ldtm[0] - 40 if idx_ldtm else 0                                  # MMA wait + loader sit just above the LDTM
stream_votes = [k for k in idx_vote if idx_ldg and k < idx_ldg[0]] or idx_vote
s1 = idx_ldg[0] - 60 if idx_ldg else (stream_votes[-1] + 600)           # first re-rank row load
after = [k for k in idx_atom if k > s1]
s2 = after[-1] if after else len(data)                                     # fail-list atomic = end of the certificate
regions = [('setup+MMA issue', 0, s0), ('stream (filter+flush)', s0, s1), ('rerank+cert', s1, s2), ('consumer', s2, len(data))]
ts = sum(int(r[si]) for r in data); ti = sum(int(r[ii]) for r in data)
print('total samples', ts, 'instructions', ti)
for name, a, b in regions:
    ch = data[a:b]
    s = sum(int(r[si]) for r in ch); i = sum(int(r[ii]) for r in ch)
    st = collections.Counter()
    for r in ch:
        for c in stall_cols:
            v = int(r[c] or 0)
            if v: st[hdr[c]] += v
    ops = collections.Counter(); opi = collections.Counter()
    for r in ch:
        o = op(r).split('.')[0]
        ops[o] += int(r[si]); opi[o] += int(r[ii])
    print(f"{name:22s} [{a}:{b}] samples {100*s/ts:5.1f}%  inst {100*i/ti:5.1f}%")
    print('    stalls:', [(k, round(100*v/ts, 1)) for k, v in st.most_common(6)])
    print('    samples by op:', [(k, round(100*v/ts, 1)) for k, v in ops.most_common(8)])
    print('    inst by op:', [(k, round(100*v/ti, 1)) for k, v in opi.most_common(8)])
