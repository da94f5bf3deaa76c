mkdir -p gpurun_out
(time python -m pytest tests/test_parity_bf16_gpu.py -q -s -p no:cacheprovider) > gpurun_out/r2_parity.log 2>&1
tail -30 gpurun_out/r2_parity.log | cut -c1-1500
(time python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_parity_bf16_gpu.py --durations=8) > gpurun_out/r2_pytest.log 2>&1
tail -25 gpurun_out/r2_pytest.log | cut -c1-400
(time python bench.py) > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
cat gpurun_out/r2_bench1.json | cut -c1-6000; tail -5 gpurun_out/r2_bench1.err
python scripts/step_profile.py 3 gpurun_out/r2_step_profile_warm.txt graph > /dev/null 2>&1
head -40 gpurun_out/r2_step_profile_warm.txt
