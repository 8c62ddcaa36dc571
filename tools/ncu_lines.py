"""Summarise an ncu source-page export (--page source --print-source cuda,sass --csv) per CUDA source line:
stall samples, instructions executed, top stall reasons.  usage: ncu_lines.py export.csv [min_samples]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
mins = int(sys.argv[2]) if len(sys.argv) > 2 else 5
hdr = None
out = []
tot = 0
cur_file = ''
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
    if len(r) > 10 and r[0] == 'Line No':
        hdr = r
        continue
    if hdr and len(r) > 10 and r[0] not in ('', '-'):
        d = dict(zip(hdr[4:], r[4:]))
        try:
            smp = int(d['# Samples'])
        except ValueError:
            continue
        tot += smp
        stalls = {k: int(v) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k and v not in ('', '-') and int(v) > 0}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
        out.append((smp, cur_file, r[0], r[1].strip()[:90], d.get('Instructions Executed'), top))
print('total samples', tot)
for smp, f, ln, src, ie, top in sorted(out, key=lambda x: -x[0]):
    if smp < mins: break
    print(f'{smp:6d} {100*smp/tot:5.1f}%  {f}:{ln:>4}  inst={ie:>7}  {src}   {top}')
