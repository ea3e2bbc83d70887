# usage: bash benchmarks/try_variants.sh  — runs bench.py with each prebuilt liblsq variant
L=local-search-quantization_b200
cp $L/liblsq_b200.so /tmp/orig.so
for v in $L/build/variants/*.so; do
  cp $v $L/liblsq_b200.so; echo "== $v"
  timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 0 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['qerror'])"
done
cp /tmp/orig.so $L/liblsq_b200.so
