#!/bin/bash
# One GPU-box visit: parity tests, per-stage timings of the bench scene and of config 4, one short bench line.
# usage: tools/gpu_check.sh <tag> [pytest-args...]
tag=${1:-run}; shift
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > $out/${tag}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q "$@" > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
tail -5 $out/${tag}_pytest.log
timeout 300 python tools/stage_profile.py 1000000 20 both ntu 3 > $out/${tag}_stages_1m.log 2>&1; cat $out/${tag}_stages_1m.log | tail -14
timeout 300 python tools/stage_profile.py 200000 20 both ntu 0 > $out/${tag}_stages_200k.log 2>&1; cat $out/${tag}_stages_200k.log | tail -14
timeout 600 python tools/stage_profile.py 5000000 5 both botanic 0 heavy close > $out/${tag}_stages_cfg4.log 2>&1; cat $out/${tag}_stages_cfg4.log | tail -14
timeout 600 python bench.py --steps 30 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.load(open("$out/${tag}_bench.json"))
    print("value", round(d["value"], 1), "ms", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), round(d["e2e"]["ms_per_step"], 4))
    print({k: v["us"] for k, v in d["roofline"]["stages"].items()})
except Exception as e:
    print("bench parse failed", e)
PY
