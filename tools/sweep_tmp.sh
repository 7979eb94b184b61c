cd $GRAFT_REPO_ROOT
for v in ballot ballot_atom; do
  for fam in phrases mixed zipf uniform; do
    FCX_LIBFCX=$PWD/build/variants/libfcx_$v.so python tools/sweep_run.py $fam 256 2>&1 | tail -1
  done
done
