#!/usr/bin/env python
"""Key metrics per launch from `ncu -i X.ncu-rep --page raw --csv`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Grid Size', 'gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'sm__cycles_elapsed.max']
idx = [hdr.index(w) if w in hdr else None for w in want]
for r in rows[2:]:
    print(' | '.join('%s=%s%s' % (w.split('.')[0][-28:], r[i], units[i] if units[i] not in ('', 'block') else '') for w, i in zip(want, idx) if i is not None))
