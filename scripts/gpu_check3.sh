set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 120 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_v3.json 2> gpurun_out/bench_train_v3.err; tail -c 1500 gpurun_out/bench_train_v3.json | head -c 700; tail -3 gpurun_out/bench_train_v3.err
python bench.py --mode sim --steps 120 --warmup 5 --no-cpu-baseline > gpurun_out/bench_sim_v3.json 2>&1; head -c 300 gpurun_out/bench_sim_v3.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1500 --csv --log-file gpurun_out/launches_train_v3.csv \
    python bench.py --replicas 2048 --burnin 120 --steps 121 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train_v3.log 2>&1
