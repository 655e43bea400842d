# Round-end evidence in one call (tag = first argument): tests, smoke, both bench arms, launch list, step timeline, SASS
# summary inputs, ncu --set full of the kernels this half of the round added.
tag=${1:-r04z}; out=gpurun_out
bash scripts/gpu_round_check.sh $tag
timeout 200 python scripts/step_timeline.py --config c4 2> $out/${tag}_timeline_c4.err | c++filt > $out/${tag}_timeline_c4.txt; tail -2 $out/${tag}_timeline_c4.err
echo "== ncu --set full: pool2 backward kernel (C4 eager step)"
timeout 200 ncu --set full --import-source on --clock-control none -k regex:pool2_relu -s 2 -c 1 -f -o /tmp/${tag}_pool2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --mode eager > $out/${tag}_ncu_pool2.log 2>&1
ncu -i /tmp/${tag}_pool2.ncu-rep --page raw --csv > $out/${tag}_pool2_full_raw.csv 2>/dev/null
echo "== ncu --set full: wide persistent conv kernel (VGG-16 eager step)"
timeout 400 ncu --set full --import-source on --clock-control none -k "regex:tc_wide|WgradProblem" -s 20 -c 10 -f -o /tmp/${tag}_wide \
    python bench.py --config c5-vgg16 --steps 1 --warmup 1 --no-cpu-baseline --no-extra --mode eager > $out/${tag}_ncu_wide.log 2>&1
ncu -i /tmp/${tag}_wide.ncu-rep --page raw --csv > $out/${tag}_wide_full_raw.csv 2>/dev/null
python scripts/ncu_top_stalls.py /tmp/${tag}_wide.ncu-rep 0 30 > $out/${tag}_wide_stalls.txt 2>&1
ls -la $out/${tag}_*
