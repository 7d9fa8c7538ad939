for d in 0 128; do
  nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader -lms 50 > gpurun_out/r3i_clk_$d.csv &
  SMI=$!
  CNL_DEBUG_EPI=$d python tools/op_times.py split 400 heatmap.block_2 > gpurun_out/r3i_ops_$d.log 2>&1
  kill $SMI
  cat gpurun_out/r3i_ops_$d.log | tail -2
  python - <<PY
rows=[l.strip().split(',') for l in open('gpurun_out/r3i_clk_$d.csv') if 'MHz' in l]
vals=[(int(r[0].split()[0]), float(r[1].split()[0])) for r in rows]
busy=[v for v in vals if v[1] > 600]
import statistics
print('debug $d: samples', len(vals), 'busy', len(busy), 'median clock', statistics.median(v[0] for v in busy) if busy else None, 'median power', statistics.median(v[1] for v in busy) if busy else None)
PY
done
