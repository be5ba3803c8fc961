set -x
mkdir -p gpurun_out
python bench.py --config rollup-main-344 --steps 3 --warmup 3 > gpurun_out/bench_344.json 2> gpurun_out/bench_344.err; tail -3 gpurun_out/bench_344.err; cat gpurun_out/bench_344.json
python bench.py --config rollup-main-2048 --steps 3 --warmup 3 > gpurun_out/bench_2048.json 2> gpurun_out/bench_2048.err; tail -3 gpurun_out/bench_2048.err; cat gpurun_out/bench_2048.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_344.csv python bench.py --config rollup-main-344 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -20 gpurun_out/launches_344.csv
ncu --set full --clock-control none --import-source on -k regex:k_slots -c 1 -o gpurun_out/prof_k_slots python bench.py --config rollup-main-344 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
