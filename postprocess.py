"""Same role as the reference's postprocess.py: read data.cfg of an output
folder, load the flow/area/pressure matrices, redimensionalise them, convert the
pressure to mmHg and plot (or, without matplotlib, save) them.

    python postprocess.py output/4cycles_last
"""
import os
import sys

import numpy as np

from arteryfe.utils import (read_output, XDMF_to_matrix, redimensionalise, unit_to_mmHg, plot_matrix)


def main(data_location):
    """data_location: the data.cfg written by ArteryNetwork.dump_metadata (a folder
    holding it is accepted too)."""
    if os.path.isdir(data_location):
        data_location = os.path.join(data_location, 'data.cfg')
    order, Nx, Nt, T0, T, L, rc, qc, rho, mesh_locations, names, locations = read_output(data_location)
    T0 = redimensionalise(rc, qc, rho, T0, 'time')
    T = redimensionalise(rc, qc, rho, T, 'time')
    t = np.linspace(T0, T, Nt)
    meshes = {int(os.path.splitext(m)[0].rsplit('_', 1)[1]): m for m in mesh_locations}
    for i, name in enumerate(names):
        for j in range(2 ** order - 1):
            if L[j] > 0.0 and j in meshes:
                M = XDMF_to_matrix(Nx, Nt, meshes[j], '%s/%s_%i.xdmf' % (locations[i], name, j), name)
                M = redimensionalise(rc, qc, rho, M, name)
                if name == 'pressure':
                    M = unit_to_mmHg(M)
                x = np.linspace(0, L[j], Nx + 1)
                try:
                    plot_matrix(t, x, M, name, '%s/%s_%i.png' % (locations[i], name, j))
                except ImportError:
                    np.save('%s/%s_%i_matrix.npy' % (locations[i], name, j), M)
                    print('matplotlib is not installed: saved %s/%s_%i_matrix.npy' % (locations[i], name, j))


if __name__ == '__main__':
    main(sys.argv[1])
