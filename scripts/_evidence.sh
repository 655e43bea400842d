# One call that refreshes the round's evidence (tag = first argument): tests, smoke, both bench arms, launch list, step timeline,
# every configuration's bench line, the ncu pass over one whole step.
tag=${1:-r02z}; out=gpurun_out
bash scripts/gpu_round_check.sh $tag
timeout 300 python scripts/step_timeline.py --config c4 2> $out/${tag}_timeline_c4.err | c++filt > $out/${tag}_timeline_c4.txt; tail -2 $out/${tag}_timeline_c4.err
for cfg in c2 c3 c5-vgg16 c5-resnet; do echo "== bench $cfg"; timeout 600 python bench.py --config $cfg > $out/${tag}_bench_$cfg.json 2> $out/${tag}_bench_$cfg.err; head -c 250 $out/${tag}_bench_$cfg.json; echo; tail -2 $out/${tag}_bench_$cfg.err; done
echo "== bench c3 with device dropout"; timeout 300 python bench.py --config c3 --dropout device --no-cpu-baseline > $out/${tag}_bench_c3_devdrop.json 2> $out/${tag}_bench_c3_devdrop.err; head -c 250 $out/${tag}_bench_c3_devdrop.json; echo
bash scripts/ncu_step_capture.sh $tag c4 900 135
