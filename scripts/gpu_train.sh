set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 120 --warmup 5 > gpurun_out/bench_train_n1.json 2> gpurun_out/bench_train_n1.err; tail -c 2500 gpurun_out/bench_train_n1.json; tail -5 gpurun_out/bench_train_n1.err
python bench.py --steps 120 --warmup 5 --tf32 --no-cpu-baseline > gpurun_out/bench_train_tf32.json 2> gpurun_out/bench_train_tf32.err; tail -c 1200 gpurun_out/bench_train_tf32.json
python bench.py --mode sim --steps 120 --warmup 5 --no-cpu-baseline > gpurun_out/bench_sim_n1.json 2>&1; tail -c 600 gpurun_out/bench_sim_n1.json
# launch list of a training segment (R reduced so that ncu's serialisation stays short): shares only
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1500 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --replicas 2048 --burnin 120 --steps 121 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train.log 2>&1
tail -3 gpurun_out/ncu_train.log
