"""Turn gpurun_out/launches_r01.csv + gpurun_out/prof_a_final.ncu-rep into the committed summaries under profiles/."""
import collections, csv, io, json, shutil, subprocess

rows = [r for r in csv.reader(open('gpurun_out/launches_r01.csv')) if r and not r[0].startswith('==')]
hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value'); iu = hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iv].replace(',', ''))
    v = v / 1e3 if r[iu] == 'ns' else (v * 1e3 if r[iu] == 'ms' else v)
    agg.setdefault(r[ik], []).append(v)
tot = sum(sum(v) for v in agg.values())
head = [k for k in agg if 'k_fused_a<' in k and ', 1, 0>(' in k]      # the MUL = 0 instances are the headline CRT / CRTInv
head_tot = sum(sum(agg[k]) for k in head)
with open('profiles/r01_launch_list.md', 'w') as f:
    f.write('# Round 1 -- ncu launch list of `python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e`\n\n')
    f.write('`ncu --metrics gpu__time_duration.sum --clock-control none -c 700` after the same command exited 0 without ncu\n')
    f.write('(cold-cache, serialised: compare shares, not absolutes).  Raw CSV: `profiles/r01_launches.csv`.\n\n')
    f.write('The timed region of the headline step launches only the two `k_fused_a` kernels (CRT, CRTInv), 65536 ring elements each;\n')
    f.write('their shares of the step under ncu: ' + ', '.join(f'{("CRTInv" if "<(bool)1" in k or "<1," in k else "CRT")} {100*sum(agg[k])/head_tot:.1f}%' for k in head) + '\n')
    bl = json.loads(open('gpurun_out/bench_r01_final.json').read().strip().splitlines()[-1])['roofline']['ms_per_launch']
    c, i = bl['tensorCRTRq'], bl['tensorCRTInvRq']
    f.write(f'(CUDA events inside bench.py, same build: CRT {c:.3f} ms, CRTInv {i:.3f} ms => {100*c/(c+i):.1f}% / {100*i/(c+i):.1f}%).  The other rows are the `per_op` and\n`other_configs` sections of bench.py and torch RNG / comparison kernels outside the timed region.\n\n')
    f.write('| kernel | launches | mean us | total us | share of all |\n|---|---|---|---|---|\n')
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f'| `{k[:120]}` | {len(v)} | {sum(v)/len(v):.1f} | {sum(v):.0f} | {100*sum(v)/tot:.1f}% |\n')
shutil.copy('gpurun_out/launches_r01.csv', 'profiles/r01_launches.csv')

raw = subprocess.run(['ncu', '-i', 'gpurun_out/prof_a_final.ncu-rep', '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max']
out = ['# Round 1 -- `ncu --set full` of the headline kernels (fused_a), batch 65536, m=14400, q=14401\n',
       'Command: `ncu --set full --clock-control none --import-source on -k regex:k_fused_a -s 6 -c 2 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-per-op`',
       '(after the same command exited 0 without ncu).  Algorithmic bytes per launch: 65536 x 61440 = 4 026 531 840.',
       'Integer-pipe utilisation asked for by BASELINE.json: `sm__inst_executed_pipe_alu` (IADD3/LOP3/SEL/VIADDMNMX) and `..._pipe_fma` (IMAD).\n']
traffic = {}
def tobytes(v, u): return float(v.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    sym = 'tensorCRTInvRq' if ('<1' in name or '(bool)1' in name) else 'tensorCRTRq'
    out.append(f'## {sym}: `{name[:110]}`\n')
    out.append('| metric | value |\n|---|---|')
    vals = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w); out.append(f'| {w} [{units[i]}] | {r[i]} |'); vals[w] = (r[i], units[i])
    rd = tobytes(*vals['dram__bytes_read.sum']); wr = tobytes(*vals['dram__bytes_write.sum'])
    traffic[sym] = rd + wr
    items = []
    for i, h in enumerate(hdr):
        if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
            try: items.append((float(r[i]), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError: pass
    items.sort(reverse=True); t = sum(v for v, _ in items) or 1
    out.append(f'| dram traffic per launch (read+write) | {rd+wr:.4g} B = {(rd+wr)/4026531840:.3f} x algorithmic |')
    out.append(f'| warp-instructions per ring element | {float(vals["smsp__inst_executed.sum"][0])/65536:.0f} |')
    out.append('| warp stall samples | ' + ', '.join(f'{h} {100*v/t:.0f}%' for v, h in items[:8]) + ' |\n')
open('profiles/r01_fused_a_ncu.md', 'w').write('\n'.join(out) + '\n')
json.dump({**traffic, 'source': 'profiles/r01_fused_a_ncu.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch, batch 65536)'},
          open('profiles/traffic.json', 'w'), indent=1)
print(open('profiles/r01_launch_list.md').read()[:3000]); print('\n'.join(out)[:2500])
