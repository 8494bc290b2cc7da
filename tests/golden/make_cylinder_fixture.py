"""Generates tests/golden/cylinder_cavity_pec.npz from the REFERENCE's own example and regression data (run in the build
container only; needs /root/reference):
  * the example mesh examples/cylinder/mesh/cylinder_hex.msh (80 HEX27 + 72 QUAD9, Gmsh 2.2 binary) converted to corner
    connectivity + lexicographic order-2 element nodes (palace_b200/host/gmsh.py),
  * the eigenfrequencies the reference's regression suite stores for examples/cylinder/cavity_pec.json (Order 4, PEC walls,
    eps_r = 2.08, tan delta = 4e-4): test/data/regression/ref/cylinder/cavity_pec/eig.csv, compared there at rtol 1e-4
    (test/unit/regression/cases.cpp:219-228).
These are END-TO-END golden numbers produced by the reference itself; tests/test_cylinder_golden.py holds the oracle's
discretisation to them."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from palace_b200.host import gmsh  # noqa: E402

REF = "/root/reference"


def main():
    m = gmsh.load_hex27(os.path.join(REF, "examples/cylinder/mesh/cylinder_hex.msh"))
    rows = []
    with open(os.path.join(REF, "test/data/regression/ref/cylinder/cavity_pec/eig.csv")) as f:
        next(f)
        for line in f:
            rows.append([float(x) for x in line.split(",")])
    rows = np.array(rows)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cylinder_cavity_pec.npz")
    np.savez_compressed(path, elems=m.elems, xe2=m.xe2, attr=m.attr, bdr_attr=m.bdr_attr, bdr_verts=m.bdr_verts, verts=m.verts,
                        ref_f_re_ghz=rows[:, 1], ref_f_im_ghz=rows[:, 2], ref_Q=rows[:, 3],
                        order=4, eps_r=2.08, loss_tan=4.0e-4, L0=1.0e-2, target_ghz=2.0)
    print("wrote", path, os.path.getsize(path), "bytes;", m.ne, "hexes,", rows.shape[0], "reference modes")


if __name__ == "__main__":
    main()
