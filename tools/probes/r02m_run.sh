# round-2 evidence run (final kernels): full GPU test suite, bench line, ncu launch list of one timed step, ncu --set full of
# conv launches of a step (summarised ON THE BOX), compute-sanitizer memcheck / racecheck over every conv kernel variant
timeout 900 python -m pytest tests -q -x -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err
tail -2 gpurun_out/r02m_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 1060 -c 345 --csv --log-file gpurun_out/r02m_launches.csv python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r02m_ncu_bench.log 2>&1
ncu --set full --clock-control none -k regex:conv_ -s 1100 -c 60 -o /tmp/r02m_conv --force-overwrite python bench.py --steps 2 --warmup 3 --no-train --no-cpu-baseline > gpurun_out/r02m_ncu_full.log 2>&1
python tools/ncu_summary.py /tmp/r02m_conv.ncu-rep gpurun_out/r02m_ncu_conv_fp16x2.md "fp16x2 mode, batch 256: 60 conv launches of a timed step (ncu --set full)" > /dev/null 2>&1
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r02m_memcheck.log python tools/probes/sanitize_target.py > gpurun_out/r02m_memcheck.out 2>&1
tail -2 gpurun_out/r02m_memcheck.out; tail -3 gpurun_out/r02m_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --log-file gpurun_out/r02m_racecheck.log python tools/probes/sanitize_target.py convs > gpurun_out/r02m_racecheck.out 2>&1
tail -2 gpurun_out/r02m_racecheck.out; tail -3 gpurun_out/r02m_racecheck.log
ls -la gpurun_out/ | tail -8
