set -x
timeout 300 python scripts/gpu_probe16.py 4096 2>&1 | grep gflops > gpurun_out/probe16.log
timeout 600 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; tail -c 300 gpurun_out/bench_final2.err
cd fault-tolerant-sgemm-on-nvidia-gpus_b200 && (timeout 900 ./ft_sgemm 1024 16384 1024 0 32 > ../gpurun_out/cli_sweep_final.txt 2> ../gpurun_out/cli_sweep_final.err); cd ..
tail -22 gpurun_out/cli_sweep_final.txt
