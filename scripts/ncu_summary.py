"""Prints the handful of ncu raw-page metrics used in profiles/ summaries.
usage: python scripts/ncu_summary.py report.ncu-rep [metric-substring ...]"""
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu.sum', 'smsp__inst_executed_op_shfl.sum',
        'l1tex__lsuin_requests.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__f_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_requests_pipe_lsu.sum', 'l1tex__t_sectors_pipe_lsu.sum']


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for d in data:
        print('---', d[idx['Kernel Name']][:70], d[idx['Grid Size']])
        for w in WANT:
            if w in idx:
                print('  ', w, d[idx[w]], units[idx[w]])
        for e in extra:
            for h in hdr:
                if e in h and h not in WANT:
                    print('  ', h, d[idx[h]], units[idx[h]])


if __name__ == '__main__':
    main()
