#!/bin/bash
# device-resident and end-to-end extraction throughput over (streams, batch); prints one line per point
# usage: sweep_streams.sh ["2 3 4 6"] ["32 64 128"]
STREAMS=${1:-"2 3 4 6"}; BATCHES=${2:-"32 64 128"}
for st in $STREAMS; do for b in $BATCHES; do
  timeout 120 python bench.py --steps 100 --warmup 5 --no-cpu --no-matching --streams $st --batch $b 2>/dev/null |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('streams', $st, 'batch', $b, 'device', round(d['value']), 'e2e', round(d['e2e']['value']))"
done; done
