mkdir -p gpurun_out
./profiles/microbench/mulmod_bench | tee gpurun_out/mulmod_bench.txt
