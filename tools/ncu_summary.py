"""Summarise an `ncu --set full` report (read here, no GPU needed) into the JSON / text files kept under profiles/.

    python tools/ncu_summary.py gpurun_out/x/ncu_fused_n3000.ncu-rep profiles/r02_ncu_relation_fused_n3000

writes <out>.json (one dict: the headline metrics of the FIRST captured launch, byte counts in bytes, times in us) and
<out>.txt (opcode mix, stall-reason mix and the 25 most-sampled SASS instructions from the source page)."""
import csv
import io
import json
import subprocess
import sys
from collections import Counter

WANT = ['launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem']
SCALE = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'us': 1, 'ns': 1e-3, 'ms': 1e3, 's': 1e6}


def ncu(rep, *args):
    return subprocess.run(['ncu', '-i', rep] + list(args), capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(io.StringIO(ncu(rep, '--page', 'raw', '--csv'))))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {'kernel': vals[hdr.index('Kernel Name')], 'report': rep, 'launches_in_report': len(rows) - 2}
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            try:
                v = float(vals[i].replace(',', ''))
            except ValueError:
                continue
            d[w] = v * SCALE.get(units[i], 1)
    json.dump(d, open(out + '.json', 'w'), indent=1)
    src = list(csv.reader(io.StringIO(ncu(rep, '--page', 'source', '--csv'))))
    h = src[1]
    ix = {k: i for i, k in enumerate(h)}
    ins, smp = ix['Instructions Executed'], ix['# Samples']
    body = [r for r in src[2:] if len(r) >= len(h) and r[ins].isdigit()]
    tot = sum(int(r[ins]) for r in body) or 1
    ts = sum(int(r[smp]) for r in body if r[smp].isdigit()) or 1
    ops, ops_s = Counter(), Counter()
    for r in body:
        t = r[1].split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        ops[op] += int(r[ins]); ops_s[op] += int(r[smp]) if r[smp].isdigit() else 0
    stall = Counter()
    for k in h:
        if k.startswith('stall_') and 'Not Issued' not in k:
            stall[k] = sum(int(r[ix[k]]) for r in body if r[ix[k]].isdigit())
    ssum = sum(stall.values()) or 1
    with open(out + '.txt', 'w') as f:
        f.write('%s\n%s\nwarp instructions executed: %d, pc samples: %d\n\nopcode mix (%% of warp instructions | %% of samples)\n' % (rep, d['kernel'], tot, ts))
        for op, n in ops.most_common(28):
            f.write('  %-10s %6.2f%% %6.2f%%\n' % (op, 100.0 * n / tot, 100.0 * ops_s[op] / ts))
        f.write('\nstall reasons (all samples)\n')
        for k, v in stall.most_common():
            f.write('  %-26s %6.2f%%\n' % (k, 100.0 * v / ssum))
        f.write('\nmost-sampled SASS instructions (samples, executed, instruction)\n')
        for r in sorted(body, key=lambda r: -(int(r[smp]) if r[smp].isdigit() else 0))[:25]:
            f.write('  %6s %10s  %s\n' % (r[smp], r[ins], r[1][:110]))
    print(json.dumps(d))


if __name__ == '__main__':
    main()
