# round-2 final check of the last build: full GPU test suite, smoke(), bench line, launch list of one step
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02o_bench.json 2> gpurun_out/r02o_bench.err
tail -2 gpurun_out/r02o_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 1060 -c 345 --csv --log-file gpurun_out/r02o_launches.csv python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r02o_ncu_bench.log 2>&1
ls -la gpurun_out | tail -4
