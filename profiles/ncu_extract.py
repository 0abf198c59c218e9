"""Print selected metrics from an `ncu --page raw --csv` dump (one column per captured launch)."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_registers', 'launch__waves_per_multiprocessor',
        'lts__t_bytes.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_cbu.sum', 'sm__inst_executed_pipe_adu.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_global_ld.sum']


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"{w:78s} {units[i]:10s}", [d[i] for d in data])
    for i, h in enumerate(hdr):
        if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct'):
            vals = [d[i] for d in data]
            try:
                if max(float(v) for v in vals) >= 3.0:
                    print(f"{h:78s} {units[i]:10s}", vals)
            except ValueError:
                pass


if __name__ == '__main__':
    main(sys.argv[1])
