"""Per-region split of an ncu source-page CSV (`ncu -i X.ncu-rep --page source --csv`) of knn_tc4_kernel.

Regions are delimited by SASS landmarks: USETMAXREG.DEALLOC (producer warps start), USETMAXREG.TRY_ALLOC (filter
warpgroups start = stream phase), the group barriers `BAR.SYNC R, 0x80` (end of stream / end of the membership
phase) and the CTA-wide barrier before the TMEM release.  Percentages are of all warp samples; the four producer
warps of a CTA are 20 % of its warps and wait at the final barrier while the filters finish ('tail')."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
data = rows[2:]
si = hdr.index('# Samples')
ii = hdr.index('Instructions Executed')
src = hdr.index('Source')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]


def op(r):
    t = r[src].split()
    return t[1] if t[0].startswith('@') else t[0]


dealloc = [k for k, r in enumerate(data) if op(r).startswith('USETMAXREG.DEALLOC')][0]
alloc = [k for k, r in enumerate(data) if op(r).startswith('USETMAXREG.TRY_ALLOC')][0]
gbar = [k for k, r in enumerate(data) if op(r).startswith('BAR.SYNC') and '0x80' in r[src]]
cbar = [k for k, r in enumerate(data) if op(r).startswith('BAR.SYNC') and '0x80' not in r[src]]
end_stream, end_member = gbar[0], gbar[1]
final = cbar[-1]
regions = [('setup', 0, dealloc), ('producer warps (TMA + tcgen05 issue)', dealloc, alloc),
           ('stream (filter + network flush)', alloc, end_stream),
           ('membership (interval test + band chains)', end_stream, end_member),
           ('consumer (+ ordered re-rank code, unused here)', end_member, final),
           ('tail (producer warps at the final barrier)', final, len(data))]
ts = sum(int(r[si]) for r in data)
ti = sum(int(r[ii]) for r in data)
print('total samples', ts, 'instructions', ti)
for name, a, b in regions:
    ch = data[a:b]
    s = sum(int(r[si]) for r in ch)
    i = sum(int(r[ii]) for r in ch)
    st = collections.Counter()
    ops = collections.Counter()
    opi = collections.Counter()
    for r in ch:
        for c in stall_cols:
            v = int(r[c] or 0)
            if v:
                st[hdr[c]] += v
        o = op(r).split('.')[0]
        ops[o] += int(r[si])
        opi[o] += int(r[ii])
    print(f"{name:48s} [{a}:{b}] samples {100 * s / ts:5.1f}%  inst {100 * i / ti:5.1f}%")
    print('    stalls:', [(k, round(100 * v / ts, 1)) for k, v in st.most_common(6)])
    print('    samples by op:', [(k, round(100 * v / ts, 1)) for k, v in ops.most_common(8)])
    print('    inst by op:', [(k, round(100 * v / ti, 1)) for k, v in opi.most_common(8)])
