set -x
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/clocks_peak.csv &
SMI=$!
./tools/fp64_peak > gpurun_out/fp64_peak_r02.json 2> gpurun_out/fp64_peak.err
kill $SMI
cat gpurun_out/fp64_peak_r02.json
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_snap.json 2> gpurun_out/bench_snap.err; tail -c 1500 gpurun_out/bench_snap.json
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_snap.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu1.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --burnin 5 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_group_fused -s 6 -c 1 -o gpurun_out/prof_fused python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --burnin 5 > gpurun_out/ncu2.log 2>&1
