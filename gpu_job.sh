set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py --config rollup-main-2048 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2048.json 2> gpurun_out/bench_2048.err; tail -3 gpurun_out/bench_2048.err; cat gpurun_out/bench_2048.json
