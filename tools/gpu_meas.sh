#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python tools/swarm_diag.py --steps 6 --dense > gpurun_out/swarm_dense.jsonl 2> gpurun_out/swarm_dense.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['e2e']['value'], json.dumps(d['stress']))"
tail -1 gpurun_out/swarm_dense.jsonl; tail -3 gpurun_out/swarm_dense.err; tail -3 gpurun_out/bench.err
