set -x
python bench.py --steps 120 --warmup 5 > gpurun_out/bench_train_v4.json 2> gpurun_out/bench_train_v4.err; head -c 400 gpurun_out/bench_train_v4.json; tail -3 gpurun_out/bench_train_v4.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 1200 --csv --log-file gpurun_out/launches_train_v4.csv \
    python bench.py --replicas 2048 --burnin 120 --steps 121 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train_v4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:policy_step_tc -s 300 -c 1 -o gpurun_out/prof_policy_tc \
    python bench.py --steps 4 --warmup 3 --burnin 240 --no-cpu-baseline > gpurun_out/ncu_policy_full.log 2>&1
tail -2 gpurun_out/ncu_policy_full.log | cut -c1-300
