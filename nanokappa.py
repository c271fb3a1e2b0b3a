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
    flag = True
    while flag:
        pop.run_timestep(geo, phonons)
        flag = (pop.current_timestep < args.iterations[0]) and not pop.finish_sim
        if max_time.total_seconds() > 0:
            flag = flag and datetime.now() - start_time < max_time
    print('Saving end of run particle data...')
    pop.write_final_state(geo)
    pop.f.close()
    pop.save_plot_real_time()
    pop.view.postprocess()
    total = datetime.now() - start_time
    n_updates = pop.current_timestep * pop.N_p
    print('---------- o ----------- o ------------- o ------------')
    print("Total time: {}  ({:.3e} particle-timestep updates/s incl. set-up and output)".format(total, n_updates / max(total.total_seconds(), 1e-9)))
    print('---------- o ----------- o ------------- o ------------')
    if output_file is not None:
        sys.stdout = sys.__stdout__
        output_file.close()
    return pop


if __name__ == '__main__':
    main()
