"""Generates tests/golden/cylinder_waveguide_tet.npz from the REFERENCE's own example and regression data (run in the build
container only; needs /root/reference):
  * examples/cylinder/mesh/cylinder_tet.msh (288 quadratic tetrahedra) with the periodic boundary pair of
    examples/cylinder/waveguide.json identified: every vertex of the receiver boundary (attribute 3) is renamed to the donor
    vertex (attribute 2) it maps to under the translation (0, 0, -5.48) -- the tetrahedral Nedelec space built on the renamed
    connectivity is then conforming across the pair, which is what MFEM's periodic mesh does (geodata.cpp periodic mapping);
  * the 15 eigenfrequencies the reference's regression suite stores for that run (ND order 4, PEC wall, eps_r = 2.08,
    tan delta = 4e-4): test/data/regression/ref/cylinder/waveguide/eig.csv, compared there at rtol 1e-4
    (test/unit/regression/cases.cpp:243-251).
These are END-TO-END golden numbers of a TETRAHEDRAL Nedelec discretisation produced by the reference itself (the
tet path had no reference-held golden below 3e-4); tests/test_cylinder_tet_golden.py holds the oracle-side tet space to them."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from palace_b200.host import gmsh  # noqa: E402

REF = "/root/reference"


def main():
    m = gmsh.load_tets(os.path.join(REF, "examples/cylinder/mesh/cylinder_tet.msh"))
    don = np.unique(m.bdr_verts[m.bdr_attr == 2])
    rec = np.unique(m.bdr_verts[m.bdr_attr == 3])
    t = np.array([0.0, 0.0, -5.48])
    ren = np.arange(m.verts.shape[0])
    for r in rec:
        d = np.minimum(np.linalg.norm(m.verts[don] - (m.verts[r] - t)[None], axis=1), np.linalg.norm(m.verts[don] - (m.verts[r] + t)[None], axis=1))
        j = int(d.argmin())
        assert d[j] < 1e-8, d[j]
        ren[r] = don[j]
    rows = []
    with open(os.path.join(REF, "test/data/regression/ref/cylinder/waveguide/eig.csv")) as f:
        next(f)
        for line in f:
            rows.append([float(x) for x in line.split(",")])
    rows = np.array(rows)
    # the same mesh, space and materials with a Floquet wave vector (examples/cylinder/floquet.json: "FloquetWaveVector": [0, 0, 0.4]):
    # test/data/regression/ref/cylinder/floquet/eig.csv -- the end-to-end golden of the periodic (mixed curl) terms
    frows = []
    with open(os.path.join(REF, "test/data/regression/ref/cylinder/floquet/eig.csv")) as f:
        next(f)
        for line in f:
            frows.append([float(x) for x in line.split(",")])
    frows = np.array(frows)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cylinder_waveguide_tet.npz")
    np.savez_compressed(path, verts=m.verts, elems=m.elems, elems_periodic=ren[m.elems], attr=m.attr, xe=m.xe, geom_order=m.order,
                        ref_f_re_ghz=rows[:, 1], ref_f_im_ghz=rows[:, 2], ref_Q=rows[:, 3], order=4, eps_r=2.08, loss_tan=4.0e-4,
                        L0=1.0e-2, target_ghz=2.0, floquet_wave_vector=np.array([0.0, 0.0, 0.4]), floquet_f_re_ghz=frows[:, 1],
                        floquet_f_im_ghz=frows[:, 2], floquet_Q=frows[:, 3])
    print("wrote", path, os.path.getsize(path), "bytes;", m.ne, "tets,", rows.shape[0], "reference modes")


if __name__ == "__main__":
    main()
