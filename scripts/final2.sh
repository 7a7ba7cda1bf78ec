set -x
cd fault-tolerant-sgemm-on-nvidia-gpus_b200 && (timeout 900 ./ft_sgemm 1024 16384 1024 0 32 > ../gpurun_out/cli_sweep_final.txt 2> ../gpurun_out/cli_sweep_final.err); cd ..
tail -25 gpurun_out/cli_sweep_final.txt
timeout 400 python scripts/fault_campaign.py 4096 31 24 > gpurun_out/campaign_4096.log 2>&1; tail -3 gpurun_out/campaign_4096.log
timeout 500 python scripts/fault_campaign.py 8192 31 16 > gpurun_out/campaign_8192.log 2>&1; tail -3 gpurun_out/campaign_8192.log
