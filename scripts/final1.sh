set -x
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 400 gpurun_out/bench_final.err; head -c 600 gpurun_out/bench_final.json
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2>/dev/null; head -c 300 gpurun_out/bench_ref_final.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_final.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ftsgemm_tc -s 2 -c 1 -o gpurun_out/prof_id31_4096_final -f python scripts/run_one.py 31 4096 3 > gpurun_out/ncu31.log 2>&1; tail -2 gpurun_out/ncu31.log
timeout 300 ncu --set full --clock-control none -k regex:encode_b -s 2 -c 1 -o gpurun_out/prof_enc_4096_final -f python scripts/run_one.py 31 4096 3 > gpurun_out/ncuenc.log 2>&1; tail -2 gpurun_out/ncuenc.log
