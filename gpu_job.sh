mkdir -p gpurun_out
python bench.py --config rollup-main-2048 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err
