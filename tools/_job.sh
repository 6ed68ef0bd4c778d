echo "=== CG auto"; python tools/gemm_bench.py 2>&1 | grep -E "fwd|dgrad|wgrad" | cut -c1-100
echo "=== CG=1";  XV_GEMM_CG=1 python tools/gemm_bench.py 2>&1 | grep -E "fwd|dgrad|wgrad" | cut -c1-72
python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-260
XV_GEMM_CG=1 python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-260
