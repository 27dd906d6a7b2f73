set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 120 --warmup 5 --no-cpu-baseline > gpurun_out/bench_train_v5.json 2> gpurun_out/bench_train_v5.err; head -c 300 gpurun_out/bench_train_v5.json; tail -3 gpurun_out/bench_train_v5.err
