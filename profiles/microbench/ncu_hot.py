"""Hot instructions of an ncu report's source page: python ncu_hot.py report.ncu-rep [min_pct] [kernel-index]
(runs `ncu -i ... --page source --csv --print-source sass` and prints the instructions that carry the stall samples)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
k = 0
while k < len(rows):
    if rows[k] and rows[k][0] == 'Kernel Name':
        print('==', rows[k][1][:150])
        hdr = rows[k + 1]
        ia, isrc, isamp = hdr.index('Address'), hdr.index('Source'), hdr.index('# Samples')
        iex = hdr.index('Instructions Executed')
        data = []
        k += 2
        while k < len(rows) and rows[k] and rows[k][0] != 'Kernel Name':
            r = rows[k]
            try:
                data.append((int(r[isamp]), r[ia], r[isrc].strip(), int(r[iex])))
            except (ValueError, IndexError):
                pass
            k += 1
        tot = sum(d[0] for d in data) or 1
        print('total samples', tot, ' instructions', len(data))
        a0 = int(data[0][1], 16) if data else 0
        for n, a, src, ex in data:
            if n > tot * min_pct / 100.0:
                print('%5.1f%%  +%05x  x%-9d %s' % (100.0 * n / tot, int(a, 16) - a0, ex, src[:100]))
    else:
        k += 1
