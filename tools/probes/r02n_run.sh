# round-2 final check: full GPU test suite, smoke(), bench line (both arms), launch list of one step
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
tail -2 gpurun_out/r02n_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02n_bench_ref.json 2> gpurun_out/r02n_bench_ref.err
cat gpurun_out/r02n_bench_ref.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -s 1060 -c 345 --csv --log-file gpurun_out/r02n_launches.csv python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r02n_ncu_bench.log 2>&1
ncu --set full --clock-control none -k regex:conv_tapwin -s 60 -c 8 -o /tmp/r02n_tapwin --force-overwrite python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r02n_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/r02n_tapwin.ncu-rep gpurun_out/r02n_ncu_tapwin_pair.md "fp16x2 mode, batch 256: conv_tapwin_kernel launches with CTA pairs (ncu --set full)" > /dev/null 2>&1
ls -la gpurun_out | tail -6
