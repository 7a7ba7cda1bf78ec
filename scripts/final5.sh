set -x
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err; tail -c 300 gpurun_out/bench_final3.err; head -c 300 gpurun_out/bench_final3.json
timeout 300 python bench.py --size 8192 --steps 150 --no-cpu > gpurun_out/bench_8192.json 2> gpurun_out/bench_8192.err; head -c 200 gpurun_out/bench_8192.json
