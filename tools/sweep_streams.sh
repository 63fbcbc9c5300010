#!/bin/bash
# device-resident and end-to-end extraction throughput over (streams, batch); prints one line per point
for st in 2 3 4 6; do for b in 32 64 128; do
  timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu --no-matching --streams $st --batch $b 2>/dev/null |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('streams', $st, 'batch', $b, 'device', round(d['value']), 'e2e', round(d['e2e']['value']))"
done; done
