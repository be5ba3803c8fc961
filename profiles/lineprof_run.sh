#!/bin/bash
# usage: profiles/lineprof_run.sh <report.ncu-rep> <out_prefix> <mangled_prefix>:<kernel_name> ...
# Per-function / per-source-line stall samples of kernels in an ncu report (needs -lineinfo and
# `ncu --import-source on`): nvdisasm gives the line table of every SASS instruction, ncu's source page the
# samples per instruction; profiles/lineprof.py joins the two by instruction order.
rep=$1; out=$2; shift 2
tmp=$(mktemp -d)
( cd $tmp && cuobjdump -xelf all /root/repo/circuits_b200/libhzwitness.so >/dev/null && for c in *.sm_100a.cubin; do nvdisasm -gi -c $c; done > all.sass 2>/dev/null )
for spec in "$@"; do
  m=${spec%%:*}; k=${spec##*:}
  ncu -i $rep --page source --csv --print-source sass -k $k > $tmp/src.csv 2>/dev/null
  python /root/repo/profiles/lineprof.py $tmp/all.sass $m $tmp/src.csv > ${out}_${k}.txt 2>&1
done
python /root/repo/profiles/ncu_summary.py $rep ${out}_raw.txt > /dev/null 2>&1
rm -rf $tmp
