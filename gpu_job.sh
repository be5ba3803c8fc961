set -x
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_2048.csv python bench.py --config rollup-main-2048 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
for k in k_smt k_eddsa k_front k_sha_blocks; do
timeout 1200 ncu --set full --clock-control none -k regex:$k -c 1 -o /tmp/prof_$k python bench.py --config rollup-main-2048 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$k.log 2>&1
python3 profiles/ncu_summary.py /tmp/prof_$k.ncu-rep gpurun_out/r1_ncu_${k}_cfg2048.txt > /dev/null
done
ls -la gpurun_out
