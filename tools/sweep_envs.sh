#!/bin/bash
# SARL rollout throughput vs envs per GPU (kernel-only roofline fraction)
for E in 4096 4736 8192 9472 16384; do
  python bench.py --envs $E --T 128 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 2>&1 | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['envs_per_gpu'], round(d['value']/1e9,3), round(d['roofline']['frac'],4), round(d['roofline']['kernel_ms_avg'],4))"
done
