#!/bin/bash
# Round 2, call W: tile ranges of the c5 build on side streams (issue-bound gradient of one range beside the memory-bound
# stages of another): chunks x streams, with and without the one-stage skew.
set -u
mkdir -p gpurun_out
for cfg in "1 1 0" "2 2 0" "2 2 1" "4 2 1" "4 4 0" "4 4 1" "8 4 1" "8 2 1" "16 4 1"; do
  set -- $cfg
  GM_DTEDGE_CHUNKS=$1 GM_DTEDGE_STREAMS=$2 GM_DTEDGE_SKEW=$3 timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-iou --no-extras 2>> gpurun_out/r2w.err | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print(json.dumps({'chunks': $1, 'streams': $2, 'skew': $3, 'ms_per_step': d['ms_per_step'], 'e2e_ms': d['e2e']['ms_per_step'], 'checksum': d['config']['merged_checksum'][0]}))" >> gpurun_out/r2w.jsonl
done
cat gpurun_out/r2w.jsonl; tail -3 gpurun_out/r2w.err
