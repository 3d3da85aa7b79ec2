"""Summarise an `ncu --page source --csv` dump of ratspn_leaf_mma_kernel: stall samples per warp role
(the roles are contiguous SASS ranges delimited by marker instructions) and the hottest instructions."""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
isrc, isamp, iex = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
cols = {n: hdr.index(n) for n in stalls}
tot = sum(int(r[isamp]) for r in data)
print("total samples", tot, "instructions", len(data))
for i in sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    r = data[i]
    st = {k[6:]: int(r[v]) for k, v in cols.items() if int(r[v]) > 0}
    print(i, r[isamp], r[iex], r[isrc].strip()[:64], st)
marks = [(i, data[i][isrc].strip()[:50]) for i in range(len(data)) if re.search(r'UTCHMMA|LDTM|UBLKCP|ATOMG|NANOSLEEP|BAR.SYNC', data[i][isrc])]
print("markers:", [(i, s.split()[0] if not s.startswith('@') else s.split()[1]) for i, s in marks][:6], "...", marks[-6:])
