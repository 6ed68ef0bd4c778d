python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-260
python tools/gemm_bench.py 2>&1 | grep -E "dgrad|fwd"
