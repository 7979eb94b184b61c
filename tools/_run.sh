python -m pytest tests -x -q -m gpu 2>&1 | tail -2
for w in "zipf 64" "uniform 1024" "phrases 1024" "mixed 1024"; do
  python tools/sweep_run.py $w 10 2>&1 | tail -1
done
