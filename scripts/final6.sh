set -x
timeout 900 python bench.py > gpurun_out/bench_final4.json 2> gpurun_out/bench_final4.err; tail -c 200 gpurun_out/bench_final4.err
cd fault-tolerant-sgemm-on-nvidia-gpus_b200 && (timeout 900 ./ft_sgemm 1024 16384 1024 0 32 > ../gpurun_out/cli_sweep_final.txt 2> ../gpurun_out/cli_sweep_final.err); cd ..
tail -3 gpurun_out/cli_sweep_final.txt | cut -c1-60
