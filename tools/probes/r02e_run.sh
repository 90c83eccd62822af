# round-2 evidence run: tests of the newest kernels, bench line, ncu launch list of one timed step, ncu --set full of the
# first 36 conv launches of a step (summarised ON THE BOX: the .ncu-rep is too large to bring back), sanitizer logs
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "split_precision or graphed or exact_modes" 2>&1 | tail -4
python bench.py > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
tail -3 gpurun_out/r02e_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 1060 -c 345 --csv --log-file gpurun_out/r02e_launches.csv python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r02e_ncu_bench.log 2>&1
ncu --set full --clock-control none -k regex:conv_ -s 1100 -c 36 -o /tmp/r02e_conv --force-overwrite python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r02e_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/r02e_conv.ncu-rep gpurun_out/r02e_ncu_conv_fp16x2.md "fp16x2 mode, batch 256: first 36 conv launches of a timed step (ncu --set full)" > /dev/null 2>&1
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r02e_memcheck.log python tools/probes/sanitize_target.py > gpurun_out/r02e_memcheck.out 2>&1
tail -2 gpurun_out/r02e_memcheck.out; tail -3 gpurun_out/r02e_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --log-file gpurun_out/r02e_racecheck.log python tools/probes/sanitize_target.py convs > gpurun_out/r02e_racecheck.out 2>&1
tail -2 gpurun_out/r02e_racecheck.out; tail -3 gpurun_out/r02e_racecheck.log
ls -la gpurun_out/ | tail -12
