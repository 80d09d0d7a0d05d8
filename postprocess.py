"""Same role as the reference's postprocess.py: read data.cfg of an output
folder, load the flow/area/pressure matrices, redimensionalise them, convert the
pressure to mmHg and plot (or, without matplotlib, save) them.

    python postprocess.py output/4cycles_last
"""
import sys

import numpy as np

from arteryfe.utils import (read_output, XDMF_to_matrix, redimensionalise, unit_to_mmHg, plot_matrix)


def main(data_location):
    order, Nx, Nt, T0, T, L, rc, qc, rho, mesh_locations, names, locations = read_output(data_location + '/data.cfg')
    T0 = redimensionalise(rc, qc, rho, T0, 'time')
    T = redimensionalise(rc, qc, rho, T, 'time')
    t = np.linspace(T0, T, Nt)
    natures = {'flow': 'flow', 'area': 'area', 'pressure': 'pressure'}
    for i in range(len(mesh_locations)):
        length = redimensionalise(rc, qc, rho, L[i], 'radius')
        x = np.linspace(0, length, Nx + 1)
        for name, location in zip(names, locations):
            M = XDMF_to_matrix(Nx, Nt, mesh_locations[i], '%s/%s_%i.xdmf' % (location, name, i), name)
            M = redimensionalise(rc, qc, rho, M, natures[name])
            if name == 'pressure':
                M = unit_to_mmHg(M)
            try:
                plot_matrix(t, x, M, name, '%s/%s_%i.png' % (location, name, i))
            except ImportError:
                np.save('%s/%s_%i_matrix.npy' % (location, name, i), M)
                print('matplotlib is not installed: saved %s/%s_%i_matrix.npy' % (location, name, i))


if __name__ == '__main__':
    main(sys.argv[1])
