#!/bin/bash
# usage: profiles/summarize.sh <tag>   (reads gpurun_out/<kernel>_<tag>.ncu-rep, launches_<tag>.csv, bench_<tag>.json)
# Writes the text summaries the bench numbers come from into profiles/.
set -e
tag=$1
cd "$(dirname "$0")/.."
for rep in gpurun_out/*_${tag}.ncu-rep; do
  [ -f "$rep" ] || continue
  name=$(basename "$rep" .ncu-rep)
  ncu -i "$rep" --page details 2>/dev/null | grep -vE "^\s*(OPT|INF|WRN)|^\s*-+\s*$" > profiles/${name}_details.txt || true
  ncu -i "$rep" --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h,u=r[0],r[1]
keep=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','dram__cycles_active.avg','gpu__dram_throughput.avg.pct','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__shared_mem_per_block','launch__occupancy_limit','smsp__thread_inst_executed_per_inst_executed.ratio','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__average_warps_issue_stalled','sm__pipe_tensor_cycles_active']
for row in r[2:]:
    print('---')
    for i,k in enumerate(h):
        if any(k==x or k.startswith(x) for x in keep) and row[i]!='': print('%s [%s] = %s'%(k,u[i],row[i]))
" > profiles/${name}_raw.txt || true
done
if [ -f gpurun_out/launches_${tag}.csv ]; then
  python3 - "$tag" <<'PY'
import csv,sys,collections
tag=sys.argv[1]
rows=[r for r in csv.reader(open('gpurun_out/launches_%s.csv'%tag)) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split('(')[0]; unit=r[13]; v=float(r[14].replace(',',''))
    if unit=='ms': v*=1e3
    elif unit=='ns': v/=1e3
    elif unit=='s' or unit=='second': v*=1e6
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
tot=sum(a[1] for a in agg.values())
with open('profiles/launches_%s.txt'%tag,'w') as f:
    f.write('# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n')
    f.write('%-60s %6s %12s %7s\n'%('kernel','count','total_us','share'))
    for k,(c,t) in agg.items(): f.write('%-60s %6d %12.1f %6.1f%%\n'%(k[:60],c,t,100*t/tot))
print(open('profiles/launches_%s.txt'%tag).read())
PY
fi
[ -f gpurun_out/bench_${tag}.json ] && cp gpurun_out/bench_${tag}.json profiles/bench_${tag}.json
ls profiles
