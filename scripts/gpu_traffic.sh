# DRAM traffic of one decoder forward launch (ncu, two metrics only)
mkdir -p gpurun_out
timeout 300 python bench.py --mode render --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain_traffic.log 2>&1 &&
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:decoder_fwd -s 3 -c 1 --csv --log-file gpurun_out/traffic.csv python bench.py --mode render --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_traffic.log 2>&1
grep -E "dram__|gpu__time|lts__" gpurun_out/traffic.csv | cut -d, -f5,13-15
