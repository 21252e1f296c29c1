"""Per-instruction stall hot spots of an .ncu-rep (source page): python tools/ncu_src_hot.py file.ncu-rep [window]"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, h):
    try: return float(r[ix[h]])
    except Exception: return 0.0
W = int(sys.argv[2]) if len(sys.argv) > 2 else 128
keys = ['stall_barrier', 'stall_long_sb', 'stall_wait', 'stall_no_inst', 'stall_math', 'stall_short_sb', 'stall_lg', 'stall_membar', 'stall_selected']
tot = sum(f(r, '# Samples') for r in data)
print('instructions', len(data), 'samples', tot, {k: round(100 * sum(f(r, k) for r in data) / tot, 1) for k in keys})
for s in range(0, len(data), W):
    seg = data[s:s + W]
    S = sum(f(r, '# Samples') for r in seg)
    ex = sum(f(r, 'Instructions Executed') for r in seg)
    top = sorted(((sum(f(r, k) for r in seg), k) for k in keys), reverse=True)[:3]
    print(f'{s:5d} {100*S/tot:5.1f}% exec {ex/1e6:7.2f}M  ' + ' '.join(f'{k[6:]}={100*v/tot:.1f}' for v, k in top) + '   ' + seg[0][1][:44])
