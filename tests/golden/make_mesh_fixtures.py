"""Convert the reference's ASCII test meshes (test/meshes/pi, test/meshes/soufflet, test/meshes/pi_cavity, test/meshes/neverworld2, with their
checked-in dist_2 / dist_8 partitions) into compact .npz fixtures, so that the GPU box -- which
has no /root/reference -- can run the config-1/2 parity tests.  Run here:

    python tests/golden/make_mesh_fixtures.py

Only raw mesh DATA is stored (coordinates, connectivity, level counts, partition vectors); every
derived array is recomputed by fesom2_b200.mesh at load time.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from fesom2_b200 import mesh as M  # noqa: E402

REF = "/root/reference/test/meshes"
OUT = os.path.dirname(os.path.abspath(__file__))

for name, cyc, dists in (("pi", 360.0, (2, 8)), ("soufflet", 4.5, (2, 8)), ("pi_cavity", 360.0, (2,)),
                         ("neverworld2", 60.0, (2, 8))):      # setups/test_neverworld2/setup.yml:23 cyclic_length: 60.0
    g = M.read_fesom_mesh(os.path.join(REF, name), cyclic_length_deg=cyc)
    parts = {f"part{n}": M.read_dist(os.path.join(REF, name), n)["part"].astype(np.int8) for n in dists}
    extra = {}
    if (g.ulevels > 1).any():                      # use_cavity: cavity_elvls.out / cavity_nlvls.out
        extra = dict(ulevels=g.ulevels.astype(np.int16), ulevels_nod2D=g.ulevels_nod2D.astype(np.int16))
    np.savez_compressed(os.path.join(OUT, f"mesh_{name}.npz"), nl=g.nl, cyclic_length_deg=cyc,
                        coord_deg=np.round(g.coord_nod2D / M.RAD, 10), elem2D_nodes=g.elem2D_nodes, edges=g.edges,
                        edge_tri=g.edge_tri, nlevels=g.nlevels.astype(np.int16),
                        nlevels_nod2D=g.nlevels_nod2D.astype(np.int16), zbar=g.zbar, **extra, **parts)
    print(name, g.N, g.T, g.E, os.path.getsize(os.path.join(OUT, f"mesh_{name}.npz")))
