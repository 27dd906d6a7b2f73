set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 120 --warmup 5 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; head -c 150 gpurun_out/bench_final_n1.json
python bench.py --impl reference --steps 120 --warmup 5 > gpurun_out/bench_final_ref.json 2>&1
python bench.py --mode sim --steps 120 --warmup 5 --no-cpu-baseline > gpurun_out/bench_final_sim.json 2>&1; head -c 150 gpurun_out/bench_final_sim.json
LLR=8192 bash scripts/gpu_ll.sh > gpurun_out/ll_final.txt 2>&1; cp gpurun_out/ll.csv gpurun_out/launches_train_final.csv; head -12 gpurun_out/ll_final.txt
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 60 --csv --log-file gpurun_out/launches_sim_final.csv python bench.py --mode sim --steps 30 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
