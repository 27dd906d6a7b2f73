set -x
python bench.py --steps 120 --warmup 5 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; head -c 200 gpurun_out/bench_final_n1.json
python bench.py --impl reference --steps 120 --warmup 5 > gpurun_out/bench_final_ref.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 900 --csv --log-file gpurun_out/launches_train_final.csv \
    python bench.py --replicas 2048 --burnin 120 --steps 121 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_train_final.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:policy_step_tc2 -s 300 -c 1 -o gpurun_out/prof_policy_tc2 \
    python bench.py --steps 4 --warmup 3 --burnin 240 --no-cpu-baseline > gpurun_out/ncu_policy2_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tsc_step_kernel -s 245 -c 1 -o gpurun_out/prof_sim_v2 \
    python bench.py --mode sim --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_sim2_full.log 2>&1
