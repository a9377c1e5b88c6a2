#!/bin/bash
# one gpurun call: tests, default bench (with clocks), launch list and ncu --set full captures of the Schur kernels,
# config sweep.  Everything lands in gpurun_out/; the summaries are copied to profiles/ by hand (tools/summarize_ncu.py).
set -x
timeout 1300 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/tests_full.log 2>&1; tail -3 gpurun_out/tests_full.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.json
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --burnin 5 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --burnin 5 > gpurun_out/ncu1.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --burnin 5 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_group_schurfn|k_prior_schur" -s 15 -c 5 -o gpurun_out/prof_schur python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --burnin 5 > gpurun_out/ncu2.log 2>&1
python tools/run_configs.py --sweeps 10 > gpurun_out/configs.log 2>&1; tail -5 gpurun_out/configs.log
