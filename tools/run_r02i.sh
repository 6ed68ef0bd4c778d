O=gpurun_out/r02i; mkdir -p $O
timeout 300 python tools/gemm_selftest.py > $O/selftest.log 2>&1; tail -1 $O/selftest.log; grep -E "dgrad_bnbwd" $O/selftest.log
timeout 200 python tools/gemm_bench.py --json $O/gemm_bench.json > $O/gemm_bench.log 2>&1; grep -E "dgrad" $O/gemm_bench.log
for mk in 1024 512; do echo "== XV_FUSE_BNBWD_MINK=$mk"; XV_FUSE_BNBWD_MINK=$mk timeout 200 python bench.py --steps 200 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'], j['e2e']['value'], j['roofline']['frac'])"; done
timeout 600 python -m pytest tests/test_layerwise_backward_gpu.py tests/test_checkpoint_gpu.py tests/test_gemm_gpu.py -m gpu -q 2>&1 | tail -4
CS=$(which compute-sanitizer || echo /usr/local/cuda/bin/compute-sanitizer)
timeout 600 $CS --tool racecheck --print-limit 20 python tools/parity_debug.py 16 60 200 > $O/sanitizer_racecheck_step_v2.log 2>&1; grep -E "RACECHECK SUMMARY|Error" $O/sanitizer_racecheck_step_v2.log | head -5
timeout 600 $CS --tool racecheck --print-limit 20 python tools/gemm_selftest.py --case conv_fwd_stats_pairs > $O/sanitizer_racecheck_gemm_stats_v2.log 2>&1; grep -E "RACECHECK SUMMARY|Error" $O/sanitizer_racecheck_gemm_stats_v2.log | head -5
NCU=$(which ncu || echo /usr/local/cuda/bin/ncu)
timeout 600 $NCU --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'gemm_kernel<\(int\)0, \(int\)2>' -s 21 -c 1 -o $O/gemm_tdnn4_src -f python tools/profile_step.py 3 > $O/gemm_src.log 2>&1
$NCU -i $O/gemm_tdnn4_src.ncu-rep --page source --csv > $O/gemm_tdnn4_source.csv 2>/dev/null; wc -l $O/gemm_tdnn4_source.csv; $NCU -i $O/gemm_tdnn4_src.ncu-rep --page raw --csv 2>/dev/null | python tools/summarize_ncu_full.py /dev/stdin 2>/dev/null | head -4
rm -f $O/gemm_tdnn4_src.ncu-rep
