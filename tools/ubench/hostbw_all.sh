#!/bin/bash
# tools/ubench/hostbw_all.sh N : hostbw on N GPUs at once, for each pinning mode; prints the per-GPU lines and the sum
N=${1:-8}
for mode in 0 1 2; do
  rm -f /tmp/hostbw_*.log
  for d in $(seq 0 $((N-1))); do tools/ubench/hostbw $d $mode 3 > /tmp/hostbw_$d.log 2>&1 & done
  wait
  cat /tmp/hostbw_*.log | awk -v m=$mode '{print} /GB\/s/ {s+=$5} END {printf "mode %d: %.1f GB/s in all\n", m, s}'
done
grep -i hugepage /sys/kernel/mm/transparent_hugepage/enabled /proc/meminfo 2>/dev/null | head -5
