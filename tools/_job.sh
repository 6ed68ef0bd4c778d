python -m pytest tests -m gpu -x -q --deselect tests/test_gemm_gpu.py 2>&1 | tail -3
python bench.py --steps 200 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-260
