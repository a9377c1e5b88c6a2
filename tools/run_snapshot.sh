#!/bin/bash
# one gpurun call: tests, default bench (with clocks and the CPU reference beside it), launch list, ncu --set full captures of
# the Schur kernels (summarised ON THE BOX: the reports exceed what gpurun brings back), config sweep, strong-scaling shares
# of one GPU, one long run.  Everything lands in gpurun_out/; the summaries are copied to profiles/ by hand.
set -x
B="python bench.py --no-cpu-baseline --no-e2e"
timeout 1300 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/tests_full.log 2>&1; tail -3 gpurun_out/tests_full.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 600 gpurun_out/bench_default.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
$B --steps 3 --warmup 3 --burnin 5 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B --steps 3 --warmup 3 --burnin 5 > gpurun_out/ncu1.log 2>&1
python tools/summarize_ncu.py launches gpurun_out/launches.csv > gpurun_out/launches.txt
$B --steps 3 --warmup 3 --burnin 5 > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_group_schurfn|k_prior_schur" -s 15 -c 5 -o /tmp/prof_schur $B --steps 3 --warmup 3 --burnin 5 > gpurun_out/ncu2.log 2>&1
python tools/summarize_ncu.py report /tmp/prof_schur.ncu-rep k_prior_schur > gpurun_out/ncu_k_prior_schur.txt 2> gpurun_out/sum1.err
python tools/summarize_ncu.py report /tmp/prof_schur.ncu-rep k_group_schurfn > gpurun_out/ncu_k_group_schurfn.txt 2> gpurun_out/sum2.err
python tools/run_configs.py --sweeps 10 > gpurun_out/configs.log 2>&1; tail -5 gpurun_out/configs.log
for c in 4096 2048 1024 512; do $B --steps 20 --warmup 3 --scaling strong --chains $c 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chains on one GPU', $c, 'ms/sweep %.3f' % d['ms_per_step'], 'draws/s %.0f' % d['value'], {k: round(v['ms_per_step'], 3) for k, v in d['roofline']['per_kernel'].items()})
"; done > gpurun_out/strong_shares.txt
cat gpurun_out/strong_shares.txt
# 200 warm-up sweeps, 1000 timed
$B --steps 1000 --warmup 200 --burnin 0 > gpurun_out/bench_long.json 2> gpurun_out/bench_long.err; tail -c 300 gpurun_out/bench_long.json
