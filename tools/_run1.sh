set -x
python -m pytest tests/test_gpu_parity.py -x -q -k "flat_split or cfg2_shape or cfg4_shape or persistent_chain or fused_peer_exchange_two or odd_walker" 2>&1 | tail -15
for f in 0 1; do
  LCF_FLAT=$f LCF_DEBUG_SHAPE=1 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras 2> gpurun_out/flat${f}_sc3.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SC3 flat=$f', d['value'], d['roofline']['frac'], d['e2e']['value'])"
  LCF_FLAT=$f python bench.py --steps 10 --warmup 3 --no-cpu --no-extras --model sc4 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('SC4 flat=$f', d['value'], d['roofline']['frac'])"
  LCF_FLAT=$f python bench.py --steps 6 --warmup 3 --no-cpu --no-extras --precision fp64 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('FP64 flat=$f', d['value'], d['roofline']['frac'])"
  LCF_FLAT=$f python tools/bench_configs.py --only cfg4 --short 2>&1 | tail -2
  LCF_FLAT=$f python tools/bench_configs.py --only cfg4 --short --tune "32,16,1;16,16,1;16,8,1;8,8,1;32,8,1" 2>&1 | tail -6
done
grep -h "launch shape" gpurun_out/flat1_sc3.err | sort | uniq -c | head
