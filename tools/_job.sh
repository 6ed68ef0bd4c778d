#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/s9_pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4) > gpurun_out/s9_smoke.log
timeout 300 python bench.py > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err
timeout 200 python tools/config_bench.py > gpurun_out/s9_config_bench.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:'bn_act|stats_pool|opt_step|pack_input' -s 24 -c 14 -o /tmp/s9_hbm -f python tools/profile_step.py 3 > gpurun_out/s9_ncu.log 2>&1
ncu -i /tmp/s9_hbm.ncu-rep --page raw --csv > gpurun_out/s9_hbm_raw.csv 2>> gpurun_out/s9_ncu.log
tail -n 3 gpurun_out/s9_pytest.log; cat gpurun_out/s9_smoke.log; python -c "import json;d=json.load(open('gpurun_out/s9_bench.json'));print('bench',d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['clocks'],d['cpu_baseline']['value'])"; tail -15 gpurun_out/s9_config_bench.log | cut -c1-300
