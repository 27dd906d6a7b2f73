set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 120 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json
python bench.py --impl reference --steps 120 --warmup 5 > gpurun_out/bench_ref.json 2>&1; tail -c 1500 gpurun_out/bench_ref.json
# every launch with its device time (serialised, cold-cache: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
# the top kernel, full set, at loaded-network state (launch ~250 = after burn-in)
ncu --set full --clock-control none --import-source on -k regex:tsc_step -s 245 -c 2 -o gpurun_out/prof_sim \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
