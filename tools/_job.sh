python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/layers_bench.py 2>&1 | grep -E "bn_act|stats_pool|copy|pool_bn"
for cfg in "-DXV_POOL_ROWS=1" "-DXV_POOL_ROWS=4" "-DXV_POOL_ROWS=1 -DXV_POOL_CPT=8" "-DXV_POOL_ROWS=2 -DXV_POOL_CPT=8"; do
  echo "=== $cfg"
  XV_EXTRA_CFLAGS="$cfg" python -m tf_kaldi_speaker_b200.build 2>&1 | grep -i error
  python tools/layers_bench.py 2>&1 | grep -E "stats_pool"
done
python -m tf_kaldi_speaker_b200.build 2>&1 | grep -i error
python bench.py --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | cut -c1-260
