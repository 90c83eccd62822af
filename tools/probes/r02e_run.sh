# round-2 evidence run: bench line, ncu launch list of one timed step, ncu --set full of the first 36 conv launches of a step
python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
tail -3 gpurun_out/r02e_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 1060 -c 345 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r02e_ncu_bench.log 2>&1
ncu --set full --clock-control none -k regex:conv_ -s 1100 -c 36 -o gpurun_out/r02e_conv --force-overwrite python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r02e_ncu_full.log 2>&1
ls -la gpurun_out/ | tail -8
