set -x
cd $GRAFT_REPO_ROOT
for fam in phrases uniform zipf; do
  python tools/prof_run.py $fam 64 1 > gpurun_out/r2b_plain_$fam.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_match -c 1 -f -o gpurun_out/r2b_match_$fam python tools/prof_run.py $fam 64 1 > gpurun_out/r2b_ncu_$fam.log 2>&1
done
python tools/decode_run.py uniform 16 > gpurun_out/r2b_plain_dec.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_decode -c 1 -f -o gpurun_out/r2b_decode_uniform python tools/decode_run.py uniform 16 > gpurun_out/r2b_ncu_dec.log 2>&1
ls -la gpurun_out/
