set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py --config rollup-main-2048 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2048.json 2> gpurun_out/bench_2048.err; tail -3 gpurun_out/bench_2048.err; cat gpurun_out/bench_2048.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_smt -c 1 -o gpurun_out/prof_k_smt python bench.py --config rollup-main-344 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_smt.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_front -c 1 -o gpurun_out/prof_k_front python bench.py --config rollup-main-344 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_front.log 2>&1
ls -la gpurun_out
