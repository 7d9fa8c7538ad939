# SM clock / board power while the headline step (batch 32 @512, CUDA-graph replays of detect()) runs for ~4 s:
#   bash tools/experiments/step_clock_trace.sh   -> gpurun_out/r02_step_clocks.csv + a one-line summary
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown --format=csv,noheader -lms 50 > gpurun_out/r02_step_clocks.csv &
SMI=$!
python tools/fwd_time.py 100 > gpurun_out/r02_step_clocks_run.log 2>&1
kill $SMI
tail -1 gpurun_out/r02_step_clocks_run.log
python - <<'PY'
import statistics
rows = [l.strip().split(', ') for l in open('gpurun_out/r02_step_clocks.csv') if 'MHz' in l]
busy = [(int(r[0].split()[0]), float(r[2].split()[0]), r[3]) for r in rows if float(r[2].split()[0]) > 700]
print(f"{len(rows)} samples, {len(busy)} under load: SM clock median {statistics.median(b[0] for b in busy)} MHz (min {min(b[0] for b in busy)}, max {max(b[0] for b in busy)}; clocks.max.sm {rows[0][1]}), "
      f"power median {statistics.median(b[1] for b in busy):.0f} W (max {max(b[1] for b in busy):.0f} W), sw_power_cap active in {sum(b[2] == 'Active' for b in busy)} of {len(busy)} samples")
PY
