mkdir -p gpurun_out/r02n8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29601 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r02n8/bench_8gpu.json 2> gpurun_out/r02n8/bench_8gpu.err; echo bench2 rc=$?
python -c "import json; j=json.load(open('gpurun_out/r02n8/bench_8gpu.json')); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['config']['parallelism'])"
timeout 300 $TR --master-port 29602 tools/dist_check_sharded.py --variant multimem > gpurun_out/r02n8/n8_allreduce_check.json 2> gpurun_out/r02n8/n8_allreduce_check.err; echo check rc=$?; tail -c 600 gpurun_out/r02n8/n8_allreduce_check.json
timeout 300 $TR --master-port 29603 tools/dist_check_syncbn.py > gpurun_out/r02n8/n8_syncbn_check.json 2> gpurun_out/r02n8/n8_syncbn.err; echo syncbn rc=$?; tail -c 500 gpurun_out/r02n8/n8_syncbn_check.json
timeout 200 $TR --master-port 29606 bench.py --workload extract --gpus 8 --steps 3 --warmup 3 > gpurun_out/r02n8/bench_extract_8gpu.json 2>/dev/null; echo ex2 rc=$?; python -c "import json; j=json.load(open('gpurun_out/r02n8/bench_extract_8gpu.json')); print(j['value'], j['e2e']['value'])"
