set -x
cd $GRAFT_REPO_ROOT
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r2u_pytest.log 2>&1; echo "rc=$?" >> $O/r2u_pytest.log; tail -4 $O/r2u_pytest.log
python bench.py > $O/r2_bench_default.json 2> $O/r2u_bench_default.err; tail -c 300 $O/r2_bench_default.json; tail -3 $O/r2u_bench_default.err
python bench.py --steps 5 --warmup 3 --no-cpu --e2e-calls 1 --sustained-steps 0 > $O/r2u_plain_for_ncu.json 2> $O/r2u_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r2_launches_1e8.csv python bench.py --steps 5 --warmup 3 --no-cpu --e2e-calls 1 --sustained-steps 0 > $O/r2u_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_rare|k_step_tab|k_mode_tables' -s 12 -c 3 -o /tmp/r2_film python bench.py --particles 1e8 --steps 3 --warmup 3 --no-cpu --e2e-calls 1 --sustained-steps 0 > $O/r2u_ncu_film.log 2>&1
python profiles/ncu_summary.py /tmp/r2_film.ncu-rep > $O/r2_film_kstep_tab_krare_kmodetables_ncu_summary.txt 2>&1
python profiles/ncu_hotspots.py /tmp/r2_film.ncu-rep k_step_tab 25 > $O/r2_film_kstep_tab_hotspots.txt 2>&1
python tests/run_cli_timing.py iterations=10000 > $O/r2_cli_timing_readme_case.json 2> $O/r2u_cli.err; cat $O/r2_cli_timing_readme_case.json
