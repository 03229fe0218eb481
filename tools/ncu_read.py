"""print selected metrics of an .ncu-rep (raw page csv): python tools/ncu_read.py file.ncu-rep [regex]"""
import csv, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else
                 r'gpu__time_duration.sum|registers_per_thread|warps_active.avg.pct|sm__throughput.avg.pct|inst_executed.sum$|'
                 r'dram__bytes_(read|write).sum$|issue_stalled_.*per_warp_active.pct|shared_mem_per_block|occupancy|'
                 r'bank_conflicts|lts__t_sector_hit_rate|sm__cycles_elapsed.max|l1tex__t_sector_hit_rate|achieved_occupancy|waves_per')
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('====', r[hdr.index('Kernel Name')][:60], 'grid', r[hdr.index('Grid Size')], 'block', r[hdr.index('Block Size')])
    for h, u, v in zip(hdr, units, r):
        if pat.search(h) and v not in ('', '0', '0.00'):
            print(f'  {h} = {v} {u}')
