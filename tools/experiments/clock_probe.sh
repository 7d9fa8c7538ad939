# SM clock / board power of the dominant conv (heads.heatmap.block_2, looped 400x) under the CNL_DEBUG_EPI timing experiments:
#   0 = real stores, 64 = TMA stores issued fully out of bounds (clipped: nothing written), 128 = stores confined to a 2-image
#   (L2-resident) region, 256 = input reads confined to a 2-image region.  Results of round 2: 1477 MHz / 1792 / 1728 / 1477.
for d in 0 64 128 256; do
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
