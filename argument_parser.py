"""Command-line / parameters-file surface of Nano-kappa, kept flag-for-flag (reference
argument_parser.py:6-176): every option is a list, ``-ff file`` is read, whitespace-split and parsed
as argv, and the results folder gets an ``_N`` suffix."""
import argparse
import os
import sys

_HELP_SUPPRESS = argparse.SUPPRESS


def initialise_parser(debug_flag=False):
    p = argparse.ArgumentParser(description='Nano-kappa on B200: Monte-Carlo phonon transport, GPU particle loop.')
    a = p.add_argument
    dbg = (lambda text: text) if debug_flag else (lambda text: _HELP_SUPPRESS)
    a('--from_file', '-ff', default='', type=str, nargs=1, help='Import arguments from file.')
    a('--geometry', '-g', default=['cuboid'], type=str, nargs=1, help='Standard shape (box, cylinder, ...) or an STL file.')
    a('--dimensions', '-d', default=[10e3, 1e3, 1e3], type=float, nargs='*', help='Dimensions in angstroms.')
    a('--scale', '-s', default=[1, 1, 1], type=float, nargs=3, help='Scaling factors (x, y, z).')
    a('--geo_rotation', '-gr', default=[0, 0, 0, 'xyz'], nargs='*', help='Euler angles in degrees and their order.')
    a('--mat_rotation', '-mr', default=[], nargs='*', help='Material index, Euler angles in degrees and order.')
    a('--isotope_scat', '-is', default=[], type=int, nargs='*', help='Materials that include isotope scattering.')
    a('--particles', '-p', default=['pmps', 1], nargs=2, help='"total" N, "pmps" per mode per subvolume, or "pv" per cubic angstrom.')
    a('--timestep', '-ts', default=[1], type=float, nargs=1, help='Timestep in picoseconds.')
    a('--iterations', '-i', default=[10000], type=int, nargs=1, help='Number of timesteps.')
    a('--max_sim_time', '-mt', default=['1-00:00:00'], type=str, nargs=1, help='Maximum wall time D-HH:MM:SS.')
    a('--subvolumes', '-sv', default=[], nargs='*', help='slice N axis | grid nx ny nz | voronoi N.')
    a('--temp_dist', '-td', default=['cold'], choices=['cold', 'hot', 'linear', 'mean', 'random', 'custom'], type=str, nargs='*')
    a('--temp_interp', '-ti', default=['nearest'], choices=['nearest', 'linear', 'radial'], type=str, nargs=1)
    a('--subvol_temp', '-st', default=[], type=float, nargs='*', help='Subvolume temperatures for --temp_dist custom.')
    a('--bound_cond', '-bc', default=[], choices=['T', 'P', 'R'], type=str, nargs='*', help='T temperature, R roughness, P periodic.')
    a('--bound_pos', '-bp', default=[], nargs='*', help='relative|absolute x1 y1 z1 x2 y2 z2 ...')
    a('--bound_values', '-bv', default=[], type=float, nargs='*', help='Temperatures [K] / roughness [angstrom].')
    a('--connect_pos', '-cp', default=[], nargs='*', help='relative|absolute points on the periodic facets, paired in order.')
    a('--fig_plot', '-fp', default=[], type=str, nargs='*')
    a('--colormap', '-cm', default=['jet'], type=str, nargs=1)
    a('--theme', '-th', default=['white'], choices=['white', 'light', 'dark'], type=str, nargs=1)
    a('--n_mean', '-nm', default=[100], type=int, nargs=1, help='Datapoints (x10 iterations) in the rolling mean.')
    a('--conv_crit', '-cc', default=[0, 1], type=float, nargs=2, help='Convergence criterion and number of consecutive checks.')
    a('--mat_folder', '-mf', default=[''], type=str, nargs='*')
    a('--poscar_file', '-pf', required=True, type=str, nargs='*')
    a('--hdf_file', '-hf', required=True, type=str, nargs='*', help='phono3py hdf5, a .npz table, or synthetic:N.')
    a('--results_folder', '-rf', default=[], type=str, nargs='*')
    # debug options
    a('--part_dist', '-pd', default=['random_subvol'], type=str, nargs=1, help=dbg('random/center _ domain/subvol, or a particle_data.txt to resume from.'))
    a('--empty_subvols', '-es', default=[], type=int, nargs='*', help=dbg('Subvolumes kept empty at initialisation.'))
    a('--subvol_material', '-sm', default=[], type=int, nargs='*', help=dbg('Material index of each subvolume.'))
    a('--reference_temp', '-rt', default=['local'], nargs=1, help=dbg('Reference temperature [K] or "local".'))
    a('--reservoir_gen', '-gn', default=['constant'], choices=['fixed_rate', 'one_to_one', 'constant'], type=str, nargs='*', help=dbg('Reservoir generation mode.'))
    a('--path_points', '-pp', default=[], nargs='*', help=dbg('Points the kappa path goes through.'))
    a('--energy_normal', '-en', default=['mean'], type=str, nargs=1, help=dbg('"fixed" or "mean" energy normalisation.'))
    a('--bound_scat', '-bs', default=['velocity'], type=str, nargs='*', help=dbg('Specular model: velocity or wavevector.'))
    a('--output', '-op', default='file', type=str, nargs=1, help=dbg('"file" -> output.txt, "screen" -> terminal.'))
    return p


def read_args(debug_flag=False, argv=None):
    argv = sys.argv if argv is None else argv
    parser = initialise_parser(debug_flag)
    if ('-ff' in argv) or ('--from_file' in argv):
        flag = '-ff' if '-ff' in argv else '--from_file'
        filename = argv[argv.index(flag) + 1]
        with open(filename, 'r') as f:
            args = parser.parse_args(f.read().split())
        args.from_file = filename
    else:
        args = parser.parse_args(argv[1:])
    return args


def get_folder_index(loc):
    """Next free ``_N`` suffix; exact ``name_<int>`` siblings only (the reference matches by substring,
    argument_parser.py:170, which breaks when e.g. ``test_results`` sits next to ``test``)."""
    base, parent = os.path.basename(loc), os.path.dirname(loc)
    if not os.path.exists(parent):
        return 0
    taken = []
    for d in os.listdir(parent):
        if d.startswith(base + '_') and d[len(base) + 1:].isdigit():
            taken.append(int(d[len(base) + 1:]))
    return max(taken) + 1 if taken else 0


def generate_results_folder(args):
    if len(args.results_folder) == 0:
        args.results_folder = os.getcwd()
        return args
    loc = os.path.normpath(os.path.relpath(args.results_folder[0]))
    if not os.path.isabs(loc):
        loc = os.path.join(os.getcwd(), loc)
    i = get_folder_index(loc)
    os.makedirs(f'{loc}_{i}', exist_ok=False)
    args.results_folder = f'{loc}_{i}'
    return args
