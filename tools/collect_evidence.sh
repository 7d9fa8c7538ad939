#!/bin/bash
# Round evidence in one gpurun call (1 GPU):  gpurun --timeout 1500 -- 'bash tools/collect_evidence.sh r02'
# Everything lands in gpurun_out/<tag>_*; copy what is to be judged into profiles/ afterwards (see profiles/README.md).
tag=${1:-r02}
out=gpurun_out
M="gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_elapsed.avg.per_second"
python -m pytest tests -m gpu -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_gpu.log; tail -3 $out/${tag}_pytest_gpu.log
python -c 'import __graft_entry__ as g; g.smoke()' > $out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" >> $out/${tag}_smoke.log; tail -2 $out/${tag}_smoke.log
python bench.py > $out/${tag}_bench_c2_n1.json 2> $out/${tag}_bench_c2_n1.err
for c in 1 4 5; do python bench.py --config $c --no-fast > $out/${tag}_bench_c${c}_n1.json 2> $out/${tag}_bench_c${c}_n1.err; done
python bench.py --config 4 --topk 300 --no-fast > $out/${tag}_bench_c4topk300_n1.json 2> $out/${tag}_bench_c4topk300_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_reference_arm.err
python tools/op_times.py split 20 > $out/${tag}_op_times.log 2>&1
python tools/bench_decode.py --net > $out/${tag}_decode_bench.log 2>&1
python tools/bench_tracker_costs.py > $out/${tag}_tracker_costs.log 2>&1
python tools/bench_inference.py > $out/${tag}_inference_folder.json 2> $out/${tag}_inference_folder.err
# launch list of 2 bench steps (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:conv_tc_kernel|stem_|pool_planes|peaks_|select_gather' -s 216 -c 108 --csv --log-file $out/${tag}_bench_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-fast > /dev/null 2>&1
# per-op metrics of one eager forward + decode (54 launches; the first detect() is skipped)
ncu --metrics $M --clock-control none -k 'regex:conv_tc_kernel|stem_|pool_planes|peaks_|select_gather' -s 108 -c 54 --csv --log-file $out/${tag}_forward_per_op_ncu.csv python tools/run_forward_once.py > $out/${tag}_forward_ops.txt 2>&1
# full captures: dominant conv (CTA-pair form), row-rolling layer1 conv, decode kernels
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 50 -c 1 -f -o $out/${tag}_tower_conv python tools/run_op.py heads.heatmap.block_2 2 > $out/${tag}_tower_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 50 -c 1 -f -o $out/${tag}_layer1_rows python tools/run_op.py backbone.layer1.1.conv1 2 > $out/${tag}_layer1_rows.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:peaks_fast|select_gather' -s 8 -c 2 -f -o $out/${tag}_decode python tools/bench_decode.py --iters 1 > $out/${tag}_decode_ncu.log 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > $out/${tag}_memcheck.log 2>&1; tail -2 $out/${tag}_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > $out/${tag}_racecheck.log 2>&1; tail -2 $out/${tag}_racecheck.log
ls -la $out | grep ${tag}_ | wc -l
