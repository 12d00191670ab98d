"""Turn an ncu report into the small CSV summary kept under profiles/ (run here, no GPU needed).

    python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_ncu_<kernel>.csv
"""
import csv
import io
import subprocess
import sys

KEEP = ('Kernel Name', 'gpu__time_duration', 'dram__bytes', 'dram__throughput', 'gpu__dram_throughput', 'dram__cycles_active',
        'sm__throughput', 'sm__warps_active', 'launch__', 'smsp__inst_executed.sum', 'issue_active', 'pipe_fp64_cycles_active.avg',
        'pipe_fma_cycles_active.avg', 'lts__t_sector_hit_rate', 'l1tex__data_bank_conflicts',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.max', 'smsp__average_warp')


def main(report, out):
    raw = subprocess.run(['ncu', '-i', report, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    keep = [i for i, h in enumerate(hdr) if any(k in h for k in KEEP) and 'ops_path' not in h]
    with open(out, 'w') as fh:
        w = csv.writer(fh)
        w.writerow(['metric', 'unit'] + ['launch%d' % i for i in range(len(rows) - 2)])
        for i in keep:
            w.writerow([hdr[i], units[i]] + [r[i] for r in rows[2:]])
    for name in ('Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
                 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
                 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
                 'launch__grid_size', 'launch__block_size'):
        if name in hdr:
            i = hdr.index(name)
            print('%-62s %s %s' % (name, [r[i][:70] for r in rows[2:]], units[i]))


if __name__ == '__main__':
    main(*sys.argv[1:3])
