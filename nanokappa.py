"""nanokappa.py -- command-line driver, same contract as the reference's (nanokappa.py:1-126):

    python nanokappa.py -ff parameters.txt

builds Geometry, Phonon and Population and calls ``pop.run_timestep`` until the iteration count, the
convergence criterion or ``--max_sim_time`` stops it; the particle loop runs on the GPU."""
import os
import re
import sys
import warnings
from datetime import datetime, timedelta

from argument_parser import generate_results_folder, read_args
from nanokappa_b200.classes.Geometry import Geometry
from nanokappa_b200.classes.Phonon import Phonon
from nanokappa_b200.classes.Population import Population

debug_flag = False


def main(argv=None):
    if not debug_flag:
        warnings.simplefilter("ignore")
    print('\nNano-kappa (B200 particle loop). Running simulation, check the results folder for the current status.')
    args = read_args(debug_flag, argv)
    world, rank = int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('RANK', 0))
    if world > 1:
        # torchrun --nproc-per-node N nanokappa.py -ff parameters.txt: one rank per GPU, particles sharded.  The host
        # set-up draws random numbers (voronoi centres, reservoir counters): all ranks must draw the same ones; rank 0
        # owns the results folder, the others get a private sub-folder for their shard's dumps and keep quiet.
        import numpy as np
        import torch
        import torch.distributed as dist
        local = int(os.environ.get('LOCAL_RANK', 0))
        from nanokappa_b200.parallel import bind_to_gpu_numa
        bind_to_gpu_numa(local)
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        if rank == 0:
            args = generate_results_folder(args)
        shared = [args.results_folder, int(np.random.randint(0, 2 ** 31 - 1))]
        dist.broadcast_object_list(shared, src=0)
        np.random.seed(shared[1])
        if rank > 0:
            args.results_folder = os.path.join(shared[0], 'rank{}'.format(rank))
            os.makedirs(args.results_folder, exist_ok=True)
            sys.stdout = open(os.path.join(args.results_folder, 'output.txt'), 'a')
    else:
        args = generate_results_folder(args)
    output = args.output[0] if isinstance(args.output, (list, tuple)) else args.output
    output_file = None
    if output == 'file':
        output_file = open(os.path.join(args.results_folder, 'output.txt'), 'a')
        sys.stdout = output_file
    with open(os.path.join(args.results_folder, 'arguments.txt'), 'w') as f:
        for key, val in vars(args).items():
            if isinstance(val, str):
                f.write(f'--{key} {val}\n')
            else:
                f.write(f'--{key} ' + ''.join(f'{i} ' for i in val) + '\n')
    d, h, m, s = [int(i) for i in re.split('-|:', args.max_sim_time[0])]
    max_time = timedelta(days=d, hours=h, minutes=m, seconds=s)
    start_time = datetime.now()
    print('---------- o ----------- o ------------- o ------------')
    print("Start at: {}".format(start_time.strftime('%Y-%m-%d %H:%M:%S')))
    print(f"Simulation name: {args.results_folder}")
    print('---------- o ----------- o ------------- o ------------')
    geo = Geometry(args)
    phonons = Phonon(args, 0)
    pop = Population(args, geo, phonons)
    # the steps between two convergence rows go to the GPU as one batch and the rows are read back asynchronously
    # (Population.step_batching); NK_STEP_BATCH=0 restores one launch sequence per run_timestep call
    pop.step_batching = os.environ.get('NK_STEP_BATCH', '1') != '0'
    loop_start = datetime.now()
    step0 = pop.current_timestep
    flag = True
    while flag:
        pop.run_timestep(geo, phonons)
        flag = (pop.current_timestep < args.iterations[0]) and not pop.finish_sim       # the same on every rank
        if max_time.total_seconds() > 0:
            if world == 1:
                flag = flag and datetime.now() - start_time < max_time
            elif pop.current_timestep % 10 == 0:
                # wall clocks differ between ranks: agree on the time limit (a rank that left alone would stall the others)
                import torch
                import torch.distributed as dist
                go = torch.tensor([int(datetime.now() - start_time < max_time)], device='cuda')
                dist.all_reduce(go, op=dist.ReduceOp.MIN)
                flag = flag and bool(go.item())
    pop.engine.synchronize()
    loop_time = (datetime.now() - loop_start).total_seconds()
    print('Saving end of run particle data...')
    pop.write_final_state(geo)
    pop.f.close()
    pop.save_plot_real_time()
    pop.view.postprocess()
    total = datetime.now() - start_time
    n_updates = pop.current_timestep * pop.N_p
    print('---------- o ----------- o ------------- o ------------')
    print("Total time: {}  ({:.3e} particle-timestep updates/s incl. set-up and output)".format(total, n_updates / max(total.total_seconds(), 1e-9)))
    print("Time loop: {:.3f} s for {} timesteps  ({:.3e} particle-timestep updates/s incl. the every-10 / every-100-step output)".format(
        loop_time, pop.current_timestep - step0, (pop.current_timestep - step0) * pop.N_p / max(loop_time, 1e-9)))
    print('---------- o ----------- o ------------- o ------------')
    if output_file is not None:
        sys.stdout = sys.__stdout__
        output_file.close()
    if world > 1:
        import torch.distributed as dist
        pop.sharded.close()
        dist.destroy_process_group()
    return pop


if __name__ == '__main__':
    main()
