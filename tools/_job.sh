set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err
cat gpurun_out/bench_g.json | cut -c1-300; tail -3 gpurun_out/bench_g.err
python tools/step_kernel_times.py 20 gpurun_out/step_kernels_g.md 2>&1 | grep -v "gemm_kernel<[01]>" | head -32
