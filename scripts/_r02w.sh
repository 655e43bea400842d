out=gpurun_out; tag=${1:-r02w}
echo "== smoke"; (timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) | tee $out/${tag}_smoke.log
echo "== all gpu tests"; (timeout 600 python -m pytest tests -m gpu -q 2>&1 | grep -v "^    \|^$" | tail -30) | tee $out/${tag}_gputests.log
echo "== bench"; timeout 300 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; head -c 400 $out/${tag}_bench.json; tail -3 $out/${tag}_bench.err
for cfg in c2 c3 c5-vgg16 c5-resnet; do echo "== bench $cfg"; timeout 600 python bench.py --config $cfg > $out/${tag}_bench_$cfg.json 2> $out/${tag}_bench_$cfg.err; head -c 300 $out/${tag}_bench_$cfg.json; echo; tail -2 $out/${tag}_bench_$cfg.err; done
timeout 300 python scripts/step_timeline.py --config c4 2> $out/${tag}_timeline_c4.err | c++filt > $out/${tag}_timeline_c4.txt; tail -3 $out/${tag}_timeline_c4.err
