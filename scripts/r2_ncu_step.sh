# ncu launch list (duration + DRAM bytes per launch) of the bench command; summarise with scripts/ncu_step_dram.py
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_b_ncuref.json 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_ncu.log 2>&1
tail -2 gpurun_out/r2_ncu.log | cut -c1-300
wc -l gpurun_out/r2_launches_bench.csv
