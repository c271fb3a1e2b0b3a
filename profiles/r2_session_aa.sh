set -x
cd $GRAFT_REPO_ROOT
for b in 16 4 8 2; do
  NK_RARE_BLOCKS_PER_SM=$b python bench.py --case c1 --eta 5 --particles 2e7 --steps 30 --warmup 5 --no-cpu --sustained-steps 0 --e2e-calls 1 > gpurun_out/r2aa_c1_$b.json 2> gpurun_out/r2aa_c1_$b.err
  python -c "import json,sys; d=json.load(open('gpurun_out/r2aa_c1_$b.json')); r=d['roofline']; print('blocks/SM=$b c1', d['value'], d['ms_per_step'], r['avg_launch_ms'], r['kernel_share_of_step'])"
  NK_RARE_BLOCKS_PER_SM=$b python bench.py --particles 1e8 --steps 30 --warmup 5 --no-cpu --sustained-steps 0 --e2e-calls 1 > gpurun_out/r2aa_film_$b.json 2> gpurun_out/r2aa_film_$b.err
  python -c "import json,sys; d=json.load(open('gpurun_out/r2aa_film_$b.json')); r=d['roofline']; print('blocks/SM=$b film', d['value'], d['ms_per_step'], r['avg_launch_ms'], r['kernel_share_of_step'])"
done
