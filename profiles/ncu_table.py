"""Markdown table (one unit per column: us, MB, %) from one or more .ncu-rep files; runs wherever ncu is installed (no GPU):
   python profiles/ncu_table.py gpurun_out/a.ncu-rep [gpurun_out/b.ncu-rep ...] > profiles/ncu_r02.md
Also writes the per-kernel CSV export next to the table when --csv DIR is given (the evidence the table was made from)."""
import csv
import os
import re
import subprocess
import sys

COLS = [('gpu__time_duration.sum', 'time us', 'us'), ('dram__bytes_read.sum', 'DRAM rd MB', 'MB'),
        ('dram__bytes_write.sum', 'DRAM wr MB', 'MB'), ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM %', '%'),
        ('lts__t_bytes.sum', 'L2 MB', 'MB'), ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'tensor pipe %', '%'),
        ('sm__inst_executed.avg.per_cycle_elapsed', 'IPC', ''), ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %', '%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %', '%'), ('launch__registers_per_thread', 'regs', '')]
SCALE = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 'ns': 1e-3, 'us': 1.0,
         'ms': 1e3}


def rows_of(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:], out


def main(argv):
    csv_dir = None
    if '--csv' in argv:
        i = argv.index('--csv')
        csv_dir = argv[i + 1]
        argv = argv[:i] + argv[i + 2:]
    print('| kernel | grid x block | ' + ' | '.join(c[1] for c in COLS) + ' | CTA/SM limit (reg/smem) | top stalls |')
    print('|' + '---|' * (len(COLS) + 4))
    for path in argv:
        hdr, units, data, raw = rows_of(path)
        if csv_dir:
            os.makedirs(csv_dir, exist_ok=True)
            open(os.path.join(csv_dir, os.path.basename(path).replace('.ncu-rep', '.raw.csv')), 'w').write(raw)
        ix = {h: i for i, h in enumerate(hdr)}
        for r in data:
            name = re.sub(r'^void |rvt::|\(.*', '', r[ix['Kernel Name']])
            cells = []
            for key, _, unit in COLS:
                if key not in ix or r[ix[key]] == '':
                    cells.append('-')
                    continue
                v = float(r[ix[key]].replace(',', ''))
                u = units[ix[key]]
                if unit in ('us', 'MB'):
                    v *= SCALE.get(u, 1.0)
                cells.append(f'{v:.2f}' if unit != '' or key.endswith('elapsed') else f'{v:.0f}')
            st = []
            for k in hdr:
                if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio') and r[ix[k]]:
                    st.append((k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], float(r[ix[k]].replace(',', ''))))
            st = ', '.join(f'{k} {v:.1f}' for k, v in sorted(st, key=lambda x: -x[1])[:3])
            grid = f"{r[ix['launch__grid_size']]} x {r[ix['launch__block_size']]}" if 'launch__grid_size' in ix else '-'
            lim = f"{r[ix.get('launch__occupancy_limit_registers', 0)]}/{r[ix.get('launch__occupancy_limit_shared_mem', 0)]}"
            print(f'| {name} | {grid} | ' + ' | '.join(cells) + f' | {lim} | {st} |')


if __name__ == '__main__':
    main(sys.argv[1:])
