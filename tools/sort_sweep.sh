#!/bin/bash
# parity tests, sort-stage sweep over the depth-bucket knobs (1 M bench scene, pose 3), per-kernel times from ncu
out=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "not config4 and not sync_free_and_captured" > $out/sweep_pytest.log 2>&1; tail -3 $out/sweep_pytest.log
for split in 2047 1023; do for bucket in 512 640 896; do
  echo "split=$split bucket=$bucket $(GSR_SORT_SPLIT=$split GSR_SORT_BUCKET=$bucket python tools/stage_profile.py 1000000 20 ours ntu 3 | grep -E 'preprocess|  sort|sum stages' | tr '\n' ' ')"
done; done 2>&1 | tee $out/sort_sweep.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $out/sort_launches.csv python tools/stage_profile.py 1000000 4 ours ntu 3 > $out/sort_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/sort_launches.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-14:]:
    print(r[4][:70].ljust(70), r[8], r[7], r[-1])
PY
timeout 300 python tools/stage_profile.py 5000000 5 ours botanic 0 heavy close 2>&1 | tail -12
