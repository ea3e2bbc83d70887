timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for pp in 1 2 4 8; do
  echo "parts $pp"
  LSQ_B200_PIPELINE_PARTS=$pp timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_codes_equal_resident_codes'])"
done
