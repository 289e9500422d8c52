"""Print the roofline-relevant metrics of every launch in an `ncu --page raw --csv` dump (profiles/ summaries)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__cycles_active.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size']
extra = sys.argv[2:]
for r in data:
    print(r[idx['Kernel Name']][:60])
    for w in want + extra:
        if w in idx:
            print(f"    {w:72s} {r[idx[w]]:>16s} {units[idx[w]]}")
