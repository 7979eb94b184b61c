# round-2 profile pass: launch list of the bench command, then --set full captures of k_match per family,
# k_encode_w / parse kernels on mixed, k_decode on zipf
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2c_plain_bench.json 2> gpurun_out/r2c_plain_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c_launches_mixed1g.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2c_ncu_bench.log 2>&1
for fam in phrases uniform zipf; do
  python tools/prof_run.py $fam 64 1 > gpurun_out/r2c_plain_$fam.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_match -c 1 -f -o gpurun_out/r2c_match_$fam python tools/prof_run.py $fam 64 1 > gpurun_out/r2c_ncu_$fam.log 2>&1
done
python tools/prof_run.py mixed 192 1 > gpurun_out/r2c_plain_mixed.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:k_encode_w|k_parse|k_hist' -c 6 -f -o gpurun_out/r2c_enc_parse_mixed python tools/prof_run.py mixed 192 1 > gpurun_out/r2c_ncu_mixed.log 2>&1
ls -la gpurun_out/
