set -x
python tools/gemm_selftest.py 2>&1 | tail -4
python -m pytest tests/test_attention_gpu.py -x -q 2>&1 | tail -25
python -m pytest tests -m gpu -x -q --deselect tests/test_attention_gpu.py 2>&1 | tail -5
python bench.py --steps 100 --warmup 5 --verbose --no-cpu-baseline > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
head -12 gpurun_out/bench_b.err; cat gpurun_out/bench_b.json
