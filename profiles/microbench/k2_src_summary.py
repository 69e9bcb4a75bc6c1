"""Summarise an `ncu --page source --csv` export of pairs_kernel: sample (= warp time) and issue shares of the FADD2
loop, the rest of the kernel body and each out-of-line function; top stalled instructions outside the loop."""
import csv, re, sys, bisect
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S, A, SRC, EX = ix['# Samples'], ix['Address'], ix['Source'], ix['Instructions Executed']
tot = sum(int(r[S] or 0) for r in data)
ssum = lambda a, b: sum(int(r[S] or 0) for r in data[a:b])
isum = lambda a, b: sum(int(r[EX] or 0) for r in data[a:b])
base = int(data[0][A], 16)
tg = sorted(set((int(re.search(r'0x([0-9a-f]+)', r[SRC]).group(1), 16) - base) // 16 for r in data if 'CALL' in r[SRC]))
fadd = [i for i, r in enumerate(data) if 'FADD2' in r[SRC]]
lo, hi = fadd[0] - 12, fadd[-1] + 3
itot = isum(0, len(data))
print('kernel %s\ntotal samples %d, warp instructions %d' % (rows[0][1], tot, itot))
print('%-28s %7s %7s' % ('region', 'time %', 'issue %'))
def line(name, a, b): print('%-28s %7.2f %7.2f' % (name, 100 * ssum(a, b) / tot, 100 * isum(a, b) / itot))
line('body before FADD2 loop', 0, lo); line('FADD2 loop', lo, hi); line('body after loop', hi, tg[0])
for a, b in zip(tg, tg[1:] + [len(data)]):
    line('function @%d (%s calls)' % (a, data[a][EX]), a, b)
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
top = sorted(((int(r[S] or 0), i) for i, r in enumerate(data) if not (lo <= i < hi)), reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for s_, i in top:
    r = data[i]
    st = sorted(((float(r[ix[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print(i, '%.2f%%' % (100 * s_ / tot), r[EX], r[SRC][:64], [(h[6:], int(v)) for v, h in st])
