#!/bin/bash
# Final single-GPU validation (under gpurun): smoke, the whole -m gpu suite with the parity dump, the default bench, the
# reference arm, the extraction bench.
O=gpurun_out/r02z; mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
XV_PARITY_DUMP=$O timeout 1500 python -m pytest tests -m gpu -q -x > $O/gputests.log 2>&1; echo "tests rc=$?"; tail -4 $O/gputests.log | head -3
timeout 300 python bench.py --verbose > $O/bench_1gpu.json 2> $O/bench_1gpu_gemm_shapes.txt; echo "bench rc=$?"
python -c "import json; j=json.load(open('$O/bench_1gpu.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'], j['roofline']['step_frac'], j['gpu_launches'], j['clocks'], j['cpu_baseline']['value'])"
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_reference.json 2>/dev/null; python -c "import json; j=json.load(open('$O/bench_reference.json')); print('reference', j['value'], j['cpu_baseline']['cores'])"
timeout 200 python bench.py --workload extract --steps 3 --warmup 3 > $O/bench_extract_1gpu.json 2>/dev/null; python -c "import json; j=json.load(open('$O/bench_extract_1gpu.json')); print('extract', j['value'], j['e2e']['value'], j['roofline']['frac'])"
