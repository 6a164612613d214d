( time python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2> gpurun_out/bench_time.txt; echo "bench rc=$?" >> gpurun_out/bench_time.txt
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>> gpurun_out/bench_time.txt; echo "ref rc=$?" >> gpurun_out/bench_time.txt
REPS=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/enc_l_final.csv python tools/prof_encoder.py > gpurun_out/ncu_final.log 2>&1
REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc2_kernel<256, 1>" -s 1 -c 1 -o gpurun_out/cg2res_r02 -f python tools/prof_encoder.py > gpurun_out/ncu_cg2res.log 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/bench_time.txt
cat gpurun_out/bench_time.txt
