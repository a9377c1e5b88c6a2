#!/bin/bash
# gpurun --gpus 8: bench.py at 1/2/4/8 GPUs, weak (default: 4096 chains per GPU) and strong (4096 chains in total) scaling
set -x
for mode in weak strong; do
  for N in 1 2 4 8; do
    if [ $N = 1 ]; then
      python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline --scaling $mode > gpurun_out/scale_${mode}_$N.json 2> gpurun_out/scale_${mode}_$N.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --scaling $mode > gpurun_out/scale_${mode}_$N.json 2> gpurun_out/scale_${mode}_$N.err
    fi
    tail -c 400 gpurun_out/scale_${mode}_$N.json
  done
done
python - <<'PY'
import json
print("mode    N   chains   ms/sweep   draws/s (device)   efficiency   draws/s (e2e)   efficiency")
for mode in ("weak", "strong"):
    base = None
    for N in (1, 2, 4, 8):
        try:
            d = json.loads(open("gpurun_out/scale_%s_%d.json" % (mode, N)).read().strip().splitlines()[-1])
        except Exception as e:
            print(mode, N, "failed", e); continue
        if base is None: base = d
        print("%-7s %d  %6d   %8.3f   %14.0f   %10.3f   %14.0f   %10.3f" % (mode, N, d["config"]["chains"], d["ms_per_step"], d["value"],
              d["value"] / (N * base["value"]), d["e2e"]["value"], d["e2e"]["value"] / (N * base["e2e"]["value"])))
PY
