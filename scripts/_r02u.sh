out=gpurun_out; tag=r02u
timeout 300 python scripts/step_timeline.py --config c4 > $out/${tag}_timeline_c4.txt 2> $out/${tag}_timeline_c4.err; tail -3 $out/${tag}_timeline_c4.err; tail -25 $out/${tag}_timeline_c4.txt
DEEPFLOWS_SIDE=0 timeout 300 python scripts/step_timeline.py --config c4 > $out/${tag}_timeline_c4_noside.txt 2> $out/${tag}_timeline_c4_noside.err; tail -12 $out/${tag}_timeline_c4_noside.txt
DEEPFLOWS_SIDE=0 timeout 300 python bench.py --no-extra --no-cpu-baseline > $out/${tag}_bench_noside.json 2> $out/${tag}_bench_noside.err; head -c 300 $out/${tag}_bench_noside.json
DFB_PDL=0 timeout 300 python bench.py --no-extra --no-cpu-baseline > $out/${tag}_bench_nopdl.json 2> $out/${tag}_bench_nopdl.err; head -c 300 $out/${tag}_bench_nopdl.json
