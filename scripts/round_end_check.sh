mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest.log 2>&1; tail -4 gpurun_out/pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err; cat gpurun_out/bench1.json | cut -c1-400
python scripts/inception_bench.py 100 bf16 full 2>&1 | tail -5 | tee gpurun_out/inception_bench_final.txt
python scripts/inception_bench.py 100 fp32 2>&1 | tail -2 | tee -a gpurun_out/inception_bench_final.txt
python scripts/inception_profile.py 100 bf16 gpurun_out/inception_profile_final.txt 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:convbn_kernel --launch-skip 7 -c 1 -o gpurun_out/convbn_6e_final -f python scripts/one_convbn.py Mixed_6e.branch7x7dbl_2 > /dev/null 2>&1
ls gpurun_out | head -30
