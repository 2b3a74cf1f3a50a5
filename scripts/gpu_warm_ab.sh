for n in w0 w1 w0 w1; do
  export GO2P_LIB=$PWD/go2_onnx_controller_b200/lib/exp_$n.so
  python bench.py --rows 131072 --steps 400 --warmup 20 --no-cpu --no-b1 --no-extras --e2e-steps 1 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$n rows 131072', round(d['ms_per_step']*1e3,2),'us', d['sharding']['l2'][:30])"
  python bench.py --rows 1048576 --steps 50 --warmup 5 --no-cpu --no-b1 --no-extras --e2e-steps 1 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$n rows 1M', round(d['ms_per_step']*1e3,2),'us')"
done
