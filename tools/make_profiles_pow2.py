"""profiles/r01_fused_pow2_df_ncu.md from gpurun_out/df4_CRT.ncu-rep / df4_CRTInv.ncu-rep (ncu --set full of the config-B kernel)."""
import csv, io, subprocess
ALG = 1024 * 2097152        # algorithmic bytes per launch: 1024 elements x 16 n k
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
out = ['# Round 1 -- `ncu --set full` of the config-B kernel `k_pow2_df` (m = 2^16, four ~30-bit primes, batch 1024)\n',
       'Command: `ncu --set full --clock-control none --import-source on -k regex:k_pow2_df -s 3 -c 1 python tools/run_op.py 65536 <4 primes> 1024 <CRT|CRTInv> 2`',
       '(after the same command exited 0 without ncu: 0.587 ms / 0.602 ms per launch by CUDA events = 55.8 % / 54.4 % of the measured HBM peak).',
       f'Algorithmic bytes per launch: 1024 x 2 097 152 = {ALG}.  Unpaired schedule (the default for tupSize 4), 5 CTAs of 128 threads per SM.\n']
def tobytes(v, u): return float(v.replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
for rep, sym in (('gpurun_out/df4_CRT.ncu-rep', 'tensorCRTRq'), ('gpurun_out/df4_CRTInv.ncu-rep', 'tensorCRTInvRq')):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw))); hdr, units, r = rows[0], rows[1], rows[2]
    out.append(f'## {sym}: `{r[hdr.index("Kernel Name")][:100]}`\n'); out.append('| metric | value |\n|---|---|')
    vals = {}
    for w in want:
        if w in hdr:
            i = hdr.index(w); out.append(f'| {w} [{units[i]}] | {r[i]} |'); vals[w] = (r[i], units[i])
    rd = tobytes(*vals['dram__bytes_read.sum']); wr = tobytes(*vals['dram__bytes_write.sum'])
    out.append(f'| dram traffic per launch | read {rd:.4g} B + write {wr:.4g} B = {(rd+wr)/ALG:.3f} x algorithmic (the excess is exchange-ring write-back) |')
    out.append(f'| warp-instructions per (element, limb) | {float(vals["smsp__inst_executed.sum"][0])/4096:.0f} |')
    items = []
    for i, h in enumerate(hdr):
        if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h:
            try: items.append((float(r[i]), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError: pass
    items.sort(reverse=True); t = sum(v for v, _ in items) or 1
    out.append('| warp stall samples | ' + ', '.join(f'{h} {100*v/t:.0f}%' for v, h in items[:9]) + ' |\n')
    hot = subprocess.run(['python', 'tools/ncu_src_hot.py', rep, '192'], capture_output=True, text=True).stdout
    out.append('Stall samples along the kernel (192-instruction windows: share of samples, executed warp-instructions, top stall reasons in % of all samples):\n\n```\n' + hot + '```\n')
open('profiles/r01_fused_pow2_df_ncu.md', 'w').write('\n'.join(out) + '\n')
print('\n'.join(out)[:1500])
