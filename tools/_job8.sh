#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29740 tools/allreduce_bench.py > gpurun_out/n${N}_arbench.json 2> gpurun_out/n${N}_arbench.err
echo "arbench rc=$?"; grep '"world"' gpurun_out/n${N}_arbench.json || grep -E "Error|error" gpurun_out/n${N}_arbench.err | head -5 | cut -c1-300
MODE=$(grep '"world"' gpurun_out/n${N}_arbench.json | python -c "import json,sys;d=json.loads(sys.stdin.read())['full 39.0 MB'];print('p2p' if d['p2p_us']<d['multimem_us'] else 'multimem')" 2>/dev/null || echo multimem)
echo "mode=$MODE"
XV_AR_MODE=$MODE timeout 300 $TR --master-port 29741 bench.py --gpus $N --no-cpu-baseline --allreduce multimem > gpurun_out/n${N}f_bench_mm.json 2> gpurun_out/n${N}f_bench_mm.err
echo "bench rc=$?"
grep '"metric"' gpurun_out/n${N}f_bench_mm.json | python -c "import json,sys;d=json.loads(sys.stdin.read());print('bench',d['value'],d['ms_per_step'],d['config']['parallelism'])" || (grep -E "Error|error" gpurun_out/n${N}f_bench_mm.err | head -8 | cut -c1-300)
