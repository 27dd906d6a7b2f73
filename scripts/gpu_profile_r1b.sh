set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 120 --warmup 5 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; head -c 150 gpurun_out/bench_final_n1.json
python bench.py --steps 30 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | head -c 400
python bench.py --impl reference --steps 120 --warmup 5 > gpurun_out/bench_final_ref.json 2>&1
python bench.py --mode sim --steps 120 --warmup 5 --no-cpu-baseline > gpurun_out/bench_final_sim.json 2>&1
LLR=8192 bash scripts/gpu_ll.sh > gpurun_out/ll_final.txt 2>&1; cp gpurun_out/ll.csv gpurun_out/launches_train_final.csv
ncu --set full --clock-control none --import-source on -k regex:policy_step_tc2 -s 300 -c 1 -o gpurun_out/prof_policy_tc2 -f \
    python bench.py --steps 4 --warmup 3 --burnin 240 --no-cpu-baseline > gpurun_out/ncu_policy2_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tsc_step_kernel -s 245 -c 1 -o gpurun_out/prof_sim_v3 -f \
    python bench.py --mode sim --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_sim3_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_bwd_tc -c 1 -o gpurun_out/prof_bwd_tc -f \
    python bench.py --steps 121 --warmup 3 --burnin 0 --no-cpu-baseline --replicas 2048 > gpurun_out/ncu_bwd_full.log 2>&1
