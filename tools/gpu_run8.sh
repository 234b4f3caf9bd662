mkdir -p gpurun_out
for cb in 33554432 134217728 268435456 1073741824; do
  echo "== chunk_bytes=$cb"
  ROGTK_CHUNK_BYTES=$cb timeout 600 python bench.py --umis 400000 --steps 2 --warmup 1 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('e2e ms', d['e2e']['ms_per_step'], 'GB/s in', d['e2e']['h2d_bytes_per_step']/d['e2e']['ms_per_step']/1e6, 'value ms', d['ms_per_step'])"
done
