timeout 600 python tools/config_bench.py --json gpurun_out/config_bench.json 2>&1 | tail -25
