set -x
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
cd fault-tolerant-sgemm-on-nvidia-gpus_b200 && (timeout 900 ./ft_sgemm 1024 16384 1024 0 32 > ../gpurun_out/cli_sweep_final.txt 2> ../gpurun_out/cli_sweep_final.err); cd ..
tail -22 gpurun_out/cli_sweep_final.txt
