set -x
timeout 200 python bench.py --impl reference --steps 30 --warmup 5 > gpurun_out/bench_ref_r1d.json 2> gpurun_out/bench_ref_r1d.err
timeout 250 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_ours_r1d.json 2> gpurun_out/bench_ours_r1d.err
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1d.log 2>&1; tail -2 gpurun_out/smoke_r1d.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_r1d.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -c 12 -s 36 -o gpurun_out/prof_r1_final2 -f python tools/stage_profile.py 1000000 2 ours > gpurun_out/ncu_full_r1d.log 2>&1
tail -1 gpurun_out/ncu_full_r1d.log
