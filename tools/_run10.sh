python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
REPS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/enc_l_final.csv python tools/prof_encoder.py > gpurun_out/ncu_final.log 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
