"""Top stall-sample SASS lines of one kernel:  python tools/ncu_hot.py rep.ncu-rep <kernel-regex> [min_pct]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
minpct = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr_at = [k for k, r in enumerate(rows) if r and r[0] == 'Address']
h = rows[hdr_at[0]]; i = {n: j for j, n in enumerate(h)}
end = hdr_at[1] - 2 if len(hdr_at) > 1 else len(rows)
body = []
for r in rows[hdr_at[0] + 1:end]:
    try: body.append((int(r[i['# Samples']]), int(r[i['Instructions Executed']]), r[i['Source']].strip()))
    except (ValueError, IndexError): pass
tot = sum(b[0] for b in body); ti = sum(b[1] for b in body)
print(f"kernel {kern}: {len(body)} SASS lines, {tot} samples, {ti} warp instructions")
for k, (s, ie, src) in enumerate(body):
    if s >= tot * minpct / 100:
        print(f"{k:5d} {100 * s / tot:5.1f}% {ie:9d}  {src[:100]}")
