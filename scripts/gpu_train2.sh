set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 120 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_v2.json 2> gpurun_out/bench_train_v2.err; tail -c 1800 gpurun_out/bench_train_v2.json; tail -3 gpurun_out/bench_train_v2.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1500 --csv --log-file gpurun_out/launches_train_v2.csv \
    python bench.py --replicas 2048 --burnin 120 --steps 121 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train_v2.log 2>&1
