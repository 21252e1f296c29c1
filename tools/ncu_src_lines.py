"""Per-CUDA-source-line sample shares of an .ncu-rep: python tools/ncu_src_lines.py file.ncu-rep [top]"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# find header row (contains '# Samples')
hi = next(i for i, r in enumerate(rows) if any(c.strip() == '# Samples' for c in r))
hdr, data = rows[hi], rows[hi + 1:]
ix = {h.strip(): i for i, h in enumerate(hdr)}
def f(r, h):
    try: return float(r[ix[h]])
    except Exception: return 0.0
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
tot = sum(f(r, '# Samples') for r in data) or 1
keys = [h for h in ix if h.startswith('stall_')]
print('columns:', list(ix)[:12])
lines = sorted(data, key=lambda r: -f(r, '# Samples'))[:top]
for r in lines:
    st = sorted(((f(r, k), k) for k in keys), reverse=True)[:3]
    print(f"{100*f(r,'# Samples')/tot:5.1f}%  exec {f(r,'Instructions Executed')/1e6:7.2f}M  " + ' '.join(f'{k[6:]}={100*v/tot:.1f}' for v, k in st) + '  | ' + ' '.join(c for c in r[:3])[:110])
