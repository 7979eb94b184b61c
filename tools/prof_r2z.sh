# final round-2 profile pass (one GPU): tests, bench line, launch list, --set full captures
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1
tail -3 gpurun_out/r2z_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_reference.json 2> gpurun_out/r2z_bench_reference.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2z_plain_bench.json 2> gpurun_out/r2z_plain_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2z_launches_mixed1g.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2z_ncu_bench.log 2>&1
for fam in phrases uniform zipf; do
  python tools/prof_run.py $fam 64 1 > gpurun_out/r2z_plain_$fam.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_match -c 1 -f -o gpurun_out/r2z_match_$fam python tools/prof_run.py $fam 64 1 > gpurun_out/r2z_ncu_$fam.log 2>&1
done
python tools/prof_run.py mixed 192 1 > gpurun_out/r2z_plain_mixed.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:k_encode_w|k_parse|k_hist' -c 3 -f -o gpurun_out/r2z_enc_parse_mixed python tools/prof_run.py mixed 192 1 > gpurun_out/r2z_ncu_mixed.log 2>&1
# k_encode_w at the bench's sizes: DRAM traffic per launch (roofline.traffic)
for w in mixed:1024 uniform:1024 phrases+run:1024 zipf:64; do
  fam=${w%%:*}; mib=${w##*:}
  ncu --set full --clock-control none -k regex:k_encode_w -c 1 -f -o gpurun_out/r2z_encw_${fam}_${mib} python tools/prof_run.py $fam $mib 1 > gpurun_out/r2z_ncu_encw_$fam.log 2>&1
done
python tools/decode_run.py zipf 64 > gpurun_out/r2z_decode.log 2>&1
ls -la gpurun_out/ | head -60
