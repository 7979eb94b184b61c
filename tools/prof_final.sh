# last round-2 pass (one GPU): tests, bench line, reference arm, full-size parity, decoder profile
set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1
tail -3 gpurun_out/r2f_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2f_bench_reference.json 2> gpurun_out/r2f_bench_reference.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2f_plain_bench.json 2> gpurun_out/r2f_plain_bench.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2f_launches_mixed1g.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2f_ncu_bench.log 2>&1
python tools/decode_dev.py zipf 64 > gpurun_out/r2f_plain_dec.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:k_dec_' -c 6 -f -o gpurun_out/r2f_decode_zipf64 python tools/decode_dev.py zipf 64 > gpurun_out/r2f_ncu_dec.log 2>&1
python tools/prof_run.py mixed 192 1 > gpurun_out/r2f_plain_mixed.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:k_parse' -c 1 -f -o gpurun_out/r2f_parse_mixed python tools/prof_run.py mixed 192 1 > gpurun_out/r2f_ncu_mixed.log 2>&1
timeout 900 python tools/full_parity.py randm-1MiB zipf-64MiB phrases-1GiB uniform-1GiB mixed-2GiB > gpurun_out/r2f_full_parity.jsonl 2> gpurun_out/r2f_full_parity.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.log 2>&1
ls -la gpurun_out/ | grep r2zz
