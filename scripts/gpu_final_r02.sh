# Round-2 measurement pass (gpurun --timeout 1800 -- 'bash scripts/gpu_final_r02.sh'); outputs under gpurun_out/r02f_*
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r02f_tests.log; cat gpurun_out/r02f_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 120 --warmup 5 > gpurun_out/r02f_bench_n1_train.json 2> gpurun_out/r02f_bench.err
python bench.py > gpurun_out/r02f_bench_default.json 2>> gpurun_out/r02f_bench.err
python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_bench_driver_window.json 2>> gpurun_out/r02f_bench.err
python bench.py --scenario real_net --steps 40 --warmup 5 > gpurun_out/r02f_bench_real_net.json 2>> gpurun_out/r02f_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_reference_arm.json 2>> gpurun_out/r02f_bench.err
python bench.py --mode sim --steps 120 --warmup 5 --no-cpu-baseline > gpurun_out/r02f_bench_sim.json 2>> gpurun_out/r02f_bench.err
python bench.py --agent ia2c --policy fc --replicas 1024 --steps 120 --warmup 5 --no-cpu-baseline > gpurun_out/r02f_bench_ia2c_fc_1024.json 2>> gpurun_out/r02f_bench.err
python bench.py --agent ia2c --replicas 1024 --steps 120 --warmup 5 --no-cpu-baseline > gpurun_out/r02f_bench_ia2c_lstm_1024.json 2>> gpurun_out/r02f_bench.err
python scripts/profile_policy_phases.py > gpurun_out/r02f_policy_phases.log 2>&1
python scripts/time_dx_kernel.py > gpurun_out/r02f_dx_kernel.log 2>&1
python scripts/profile_bptt_phases.py > gpurun_out/r02f_bptt_phases.log 2>&1
TSC_BPTT_TMA=0 python scripts/profile_bptt_phases.py > gpurun_out/r02f_bptt_phases_cpasync.log 2>&1
LLR=8192 bash scripts/gpu_ll.sh > gpurun_out/r02f_launches_train_summary.txt 2>&1; cp gpurun_out/ll.csv gpurun_out/r02f_launches_train.csv
ncu --set full --clock-control none --import-source on -k regex:policy_step_tc2 -s 300 -c 1 -o gpurun_out/r02f_prof_policy -f \
    python bench.py --steps 4 --warmup 3 --burnin 240 --no-cpu-baseline > gpurun_out/r02f_ncu_policy.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm_bwd_tc_staged -c 1 -o gpurun_out/r02f_prof_bptt -f \
    python bench.py --steps 121 --warmup 3 --burnin 0 --no-cpu-baseline --replicas 2048 --chunk 1024 > gpurun_out/r02f_ncu_bptt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dx_tc_kernel -c 1 -o gpurun_out/r02f_prof_dx -f \
    python bench.py --steps 121 --warmup 3 --burnin 0 --no-cpu-baseline --replicas 2048 --chunk 1024 > gpurun_out/r02f_ncu_dx.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_tc -c 1 -o gpurun_out/r02f_prof_wgrad -f \
    python bench.py --steps 121 --warmup 3 --burnin 0 --no-cpu-baseline --replicas 2048 --chunk 1024 > gpurun_out/r02f_ncu_wgrad.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:heads_loss -c 1 -o gpurun_out/r02f_prof_heads -f \
    python bench.py --steps 121 --warmup 3 --burnin 0 --no-cpu-baseline --replicas 2048 --chunk 1024 > gpurun_out/r02f_ncu_heads.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fc_bwd_tc -c 1 -o gpurun_out/r02f_prof_fcbwd -f \
    python bench.py --steps 121 --warmup 3 --burnin 0 --no-cpu-baseline --replicas 2048 --chunk 1024 > gpurun_out/r02f_ncu_fcbwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tsc_step -s 245 -c 1 -o gpurun_out/r02f_prof_sim -f \
    python bench.py --mode sim --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_ncu_sim.log 2>&1
python - <<'PY'
import json
for f in ["n1_train","default","driver_window","real_net","reference_arm","sim","ia2c_fc_1024","ia2c_lstm_1024"]:
    try:
        d=json.loads(open("gpurun_out/r02f_bench_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.2fM"%(d["value"]/1e6), "steady", d.get("value_steady"), "ms/step %.4f"%d.get("ms_per_step",0), "e2e %.2fM"%(d["e2e"]["value"]/1e6), "roofline", d.get("roofline",{}).get("frac"))
    except Exception as e: print(f, "ERR", e)
PY
cat gpurun_out/r02f_launches_train_summary.txt | head -14
