"""Print the key metrics of every kernel in an .ncu-rep (run where ncu is installed, no GPU needed):
python profiles/ncu_summary.py gpurun_out/prof.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum']


def main(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('----', d['Kernel Name'][:70], 'grid', d.get('launch__grid_size'))
        for w in WANT:
            if w in d:
                print(f'  {w:72s} {d[w]:>16s} {units[hdr.index(w)]}')
        st = []
        for k in hdr:
            if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio') and d[k]:
                st.append((k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')], float(d[k].replace(',', ''))))
        for k, v in sorted(st, key=lambda x: -x[1])[:7]:
            print(f'     stall {k:32s} {v:8.2f} warps/issue')


if __name__ == '__main__':
    main(sys.argv[1])
