import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from palace_b200 import capi
from palace_b200.host import coeff as cf, hexmesh as hm, hexspace as hs
ctx = capi.Ctx(0); capi.set_stream(ctx)
p, n = 3, 29
mesh = hm.box_mesh(n, (1.0, 1.0, 1.0)); topo = hs.build_topology(mesh); h1 = hs.build_h1_space(mesh, topo, p)
nodes = hs.gauss_lobatto(2); xe = mesh.node_coords(1, nodes); qx, qw = hs.gauss_legendre(p + 1); nB, nG = hs.lagrange_table(nodes, qx)
geom = capi.Geom.hex(ctx, xe, mesh.attr, 1, p + 1, nB, nG, qw); t = hs.tables_1d(p, p + 1)
op = capi.Op.create(ctx, geom, capi.H1_DIFFUSION, p, h1.ndofs, h1.lex_gid.astype(np.int32), None, None, None, t.Bc, t.Gc, cf.coeff_ctx())
x = torch.rand(h1.ndofs, dtype=torch.float64, device="cuda"); y = torch.zeros_like(x)
for _ in range(6): op.apply_add(x, y)
torch.cuda.synchronize()
