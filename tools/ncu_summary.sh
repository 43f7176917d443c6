#!/bin/bash
# Key metrics of an .ncu-rep (read here, no GPU): duration, DRAM bytes, throughput %, tensor pipe %, occupancy, registers.
for f in "$@"; do
  echo "== $f"
  ncu -i "$f" --page raw --csv 2>/dev/null | python3 -c '
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]; units=rows[1]
want=["Kernel Name","Grid Size","Block Size","gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","lts__t_bytes.sum","lts__throughput.avg.pct_of_peak_sustained_elapsed","sm__throughput.avg.pct_of_peak_sustained_elapsed","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_tensor.sum","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__occupancy_limit_registers","launch__occupancy_limit_shared_mem","smsp__issue_active.avg.pct","sm__inst_executed.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","smsp__cycles_active.avg","l1tex__t_sector_hit_rate.pct","lts__t_sector_hit_rate.pct"]
idx={h:i for i,h in enumerate(hdr)}
for r in rows[2:]:
    print("; ".join(f"{w.split(chr(46))[0] if False else w}={r[idx[w]]}{units[idx[w]] if units[idx[w]] else str()}" for w in want if w in idx))
'
done
