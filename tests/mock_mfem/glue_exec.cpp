// TEST INFRASTRUCTURE ONLY. Executes the MFEM -> descriptor glue of include/b2p_palace.hpp (GatherHexNDSpace, CreateHexGeometry,
// CreateHexNDIntegrator, palace::Operator::Mult / AssembleDiagonal) on a small hexahedral mesh through the C ABI -- against
// the emulation build of the kernels on a CPU box, against libb2p.so on a GPU box -- and compares with the oracle's dense
// reference-style apply (liboracle.so). The "MFEM" is tests/mock_mfem/mfem.hpp, filled with what MFEM would hold for the mesh:
// native ND dof numbering with sign flips (-1-d), H1 nodes in MFEM's native (vertex) order, column-major DofToQuad tables.
//   glue_exec <coeff_ctx.bin> <order> [hex | dense | dense_co]     prints "GLUE_EXEC OK apply_err=... diag_err=..."
// "dense": the same mesh through the non-tensor glue (GatherDenseNDSpace, CreateGeneralGeometry, CreateDenseNDIntegrator: FULL
// DofToQuad tables, element transformations); "dense_co": with element DofTransformations (tridiagonal curl orientation).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <random>
#include <string>
#include <vector>

#include "b2p_palace.hpp"

extern "C"
{
void orc_nd_hex_1d(int p, int q1d, double *Bo, double *Bc, double *Gc, double *qw);
void orc_nd_hex_dofmap(int p, int *dof_map);
void orc_nd_hex_tables(int p, int q1d, double *interp, double *curl, double *qw);
void orc_gauss_legendre(int n, double *x, double *w);
void orc_geom_hex_qdata(int ne, int k, int q1d, const double *xe, const int *attr, double *qdata);
void orc_apply_add(int kind, int ne, int P, int Q, const double *interp, const double *deriv, const int *idx, const signed char *orient,
                   const double *qdata, const void *ctx, const double *x, double *y);
void orc_apply_add_co(int kind, int ne, int P, int Q, const double *interp, const double *deriv, const int *idx, const signed char *co,
                      const double *qdata, const void *ctx, const double *x, double *y);
void orc_diag_add(int kind, int ne, int P, int Q, const double *interp, const double *deriv, const int *idx, const double *qdata,
                  const void *ctx, double *diag);
}

static std::vector<double> col_major(const std::vector<double> &row, int nq, int nd)
{
  std::vector<double> c((size_t)nq * nd);
  for (int q = 0; q < nq; q++)
    for (int d = 0; d < nd; d++) c[q + (size_t)nq * d] = row[(size_t)q * nd + d];
  return c;
}

int main(int argc, char **argv)
{
  if (argc < 3) return 2;
  std::ifstream f(argv[1], std::ios::binary);
  std::vector<char> blob((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  const int p = std::atoi(argv[2]), q1d = p + 1, n = p + 1, P = 3 * p * n * n, Q = q1d * q1d * q1d;
  const int kind = B2P_CURLCURL_MASS;
  const std::string mode = argc > 3 ? argv[3] : "hex";

  // ---- what MFEM holds: two hexahedra side by side, trilinear nodes (slightly distorted), ND space of order p ----
  const int ne = 2;
  mfem::IntegrationRule ir;
  {
    std::vector<double> x(q1d), w(q1d);
    orc_gauss_legendre(q1d, x.data(), w.data());
    for (int i = 0; i < q1d; i++) ir.pts.push_back({x[i], w[i]});
  }
  std::vector<double> Bo(q1d * p), Bc(q1d * n), Gc(q1d * n), qw1(q1d);
  orc_nd_hex_1d(p, q1d, Bo.data(), Bc.data(), Gc.data(), qw1.data());
  mfem::VectorTensorFiniteElement nd_fe(p);
  nd_fe.closed_maps.ndof = n;
  nd_fe.closed_maps.nqpt = q1d;
  nd_fe.closed_maps.B = col_major(Bc, q1d, n);
  nd_fe.closed_maps.G = col_major(Gc, q1d, n);
  nd_fe.open_maps.ndof = p;
  nd_fe.open_maps.nqpt = q1d;
  nd_fe.open_maps.B = col_major(Bo, q1d, p);
  {
    std::vector<int> dm(P);
    orc_nd_hex_dofmap(p, dm.data());
    nd_fe.dof_map.Assign(dm);
  }
  // the two elements keep disjoint dof sets (a discontinuous space is enough for the glue: sharing is MFEM's business) and
  // every third native dof of element 1 is flipped
  mfem::FiniteElementSpace nd_fes;
  nd_fes.fe = &nd_fe;
  nd_fes.ne = ne;
  nd_fes.vsize = ne * P;
  nd_fes.elem_dofs.resize(ne);
  for (int e = 0; e < ne; e++)
    for (int i = 0; i < P; i++)
    {
      const int d = e * P + (e == 0 ? i : P - 1 - i);  // element 1 numbers its dofs backwards
      nd_fes.elem_dofs[e].push_back((e == 1 && i % 3 == 0) ? -1 - d : d);
    }
  // nodes: order-1 H1 hexahedron; MFEM's native node order = vertex order, lexicographic -> native {0,1,3,2,4,5,7,6}
  mfem::NodalTensorFiniteElement h1_fe(1);
  h1_fe.dof_map.Assign({0, 1, 3, 2, 4, 5, 7, 6});
  h1_fe.maps.ndof = 2;
  h1_fe.maps.nqpt = q1d;
  {
    std::vector<double> nB(q1d * 2), nG(q1d * 2);
    for (int q = 0; q < q1d; q++)
    {
      nB[q * 2 + 0] = 1.0 - ir.pts[q].x;
      nB[q * 2 + 1] = ir.pts[q].x;
      nG[q * 2 + 0] = -1.0;
      nG[q * 2 + 1] = 1.0;
    }
    h1_fe.maps.B = col_major(nB, q1d, 2);
    h1_fe.maps.G = col_major(nG, q1d, 2);
  }
  mfem::FiniteElementSpace node_fes;
  node_fes.fe = &h1_fe;
  node_fes.ne = ne;
  node_fes.vdim = 3;
  const int nverts = 12;  // 3 x 2 x 2 vertices
  node_fes.vsize = 3 * nverts;
  auto vid = [](int i, int j, int k) { return i + 3 * (j + 2 * k); };
  node_fes.elem_dofs.resize(ne);
  for (int e = 0; e < ne; e++)  // MFEM hexahedron vertex order
    node_fes.elem_dofs[e] = {vid(e, 0, 0), vid(e + 1, 0, 0), vid(e + 1, 1, 0), vid(e, 1, 0), vid(e, 0, 1), vid(e + 1, 0, 1), vid(e + 1, 1, 1), vid(e, 1, 1)};
  mfem::GridFunction nodes(&node_fes, 3 * nverts);
  std::mt19937 rng(7);
  std::uniform_real_distribution<double> jit(-0.08, 0.08);
  for (int k = 0; k < 2; k++)
    for (int j = 0; j < 2; j++)
      for (int i = 0; i < 3; i++)
      {
        const int v = vid(i, j, k);
        nodes[v] = 0.7 * i + jit(rng);
        nodes[nverts + v] = 0.9 * j + jit(rng);
        nodes[2 * nverts + v] = 1.1 * k + jit(rng);
      }
  mfem::Mesh mesh;
  mesh.ne = ne;
  mesh.attributes = {1, 2};
  mesh.nodes = &nodes;

  // ---- the glue under test ----
  b2p_ctx *ctx = nullptr;
  if (b2p_ctx_create(0, &ctx) != B2P_SUCCESS)
  {
    std::fprintf(stderr, "b2p_ctx_create: %s\n", b2p_last_error(nullptr));
    return 3;
  }
  // dense tables in native order, as fe.GetDofToQuad(ir, FULL) holds them, and the 3-D rule (x fastest)
  std::vector<double> interp((size_t)3 * Q * P), curl((size_t)3 * Q * P), qw(Q);
  orc_nd_hex_tables(p, q1d, interp.data(), curl.data(), qw.data());
  mfem::IntegrationRule ir3;
  for (int k = 0; k < q1d; k++)
    for (int j = 0; j < q1d; j++)
      for (int i = 0; i < q1d; i++)
        ir3.pts.push_back({ir.pts[i].x, ir.pts[i].weight * ir.pts[j].weight * ir.pts[k].weight, ir.pts[j].x, ir.pts[k].x});
  mfem::VectorFiniteElement vfe(p);
  vfe.full_maps.ndof = P;
  vfe.full_maps.nqpt = Q;
  vfe.full_maps.Bt = interp;
  vfe.full_maps.Gt = curl;
  mfem::FiniteElementSpace dense_fes = nd_fes;
  dense_fes.fe = &vfe;
  // element transformations for "dense_co": unit diagonal with sprinkled 2 x 2 blocks of small integers on consecutive dof pairs
  std::vector<mfem::DofTransformation> dts(ne);
  std::vector<signed char> co_ref;
  if (mode == "dense_co")
  {
    std::mt19937 r2(11);
    co_ref.assign((size_t)ne * P * 3, 0);
    for (int e = 0; e < ne; e++)
    {
      dts[e].n = P;
      dts[e].Minv.assign((size_t)P * P, 0.0);
      for (int i = 0; i < P; i++) dts[e].Minv[(size_t)i * P + i] = (r2() % 2) ? 1.0 : -1.0;
      for (int j = 0; j + 1 < P; j += 6)
      {
        const int b[4] = {(int)(r2() % 3) - 1, (int)(r2() % 3) - 1, (int)(r2() % 3) - 1, (int)(r2() % 3) - 1};
        dts[e].Minv[(size_t)j * P + j] = b[0];
        dts[e].Minv[(size_t)j * P + j + 1] = b[1];
        dts[e].Minv[(size_t)(j + 1) * P + j] = b[2];
        dts[e].Minv[(size_t)(j + 1) * P + j + 1] = b[3];
      }
      dense_fes.elem_trans.push_back(&dts[e]);
      // expected rows: T(r, j) = sign_j M(r, j) stored row-major tridiagonal {T(r, r-1), T(r, r), T(r, r+1)}
      for (int r = 0; r < P; r++)
        for (int t = -1; t <= 1; t++)
        {
          const int j = r + t;
          if (j < 0 || j >= P) continue;
          const double sj = nd_fes.elem_dofs[e][j] >= 0 ? 1.0 : -1.0;
          co_ref[((size_t)e * P + r) * 3 + (t + 1)] = (signed char)(sj * dts[e].Minv[(size_t)r * P + j]);
        }
    }
  }
  b2p_geom *geom = nullptr;
  auto op = std::make_unique<b2p::palace::Operator>(ctx, nd_fes.GetVSize(), nd_fes.GetVSize());
  if (mode == "hex")
  {
    geom = b2p::palace::CreateHexGeometry(ctx, mesh, ir);
    const b2p::palace::HexNDSpaceInputs in = b2p::palace::GatherHexNDSpace(nd_fes, ir);
    op->AddSubOperator(b2p::palace::CreateHexNDIntegrator(ctx, geom, kind, in, blob.data(), blob.size()));
  }
  else
  {
    geom = b2p::palace::CreateGeneralGeometry(ctx, mesh, ir3);
    const b2p::palace::DenseNDSpaceInputs in = b2p::palace::GatherDenseNDSpace(dense_fes, ir3);
    if (mode == "dense_co" && std::vector<signed char>(in.curl_orient.begin(), in.curl_orient.end()) != co_ref)
    {
      std::printf("GLUE_EXEC FAIL curl_orient rows differ from the expected tridiagonal matrices\n");
      return 1;
    }
    op->AddSubOperator(b2p::palace::CreateDenseNDIntegrator(ctx, geom, kind, in, blob.data(), blob.size()));
  }
  const int N = nd_fes.GetVSize();
  std::vector<double> x(N), y(N), d(N);
  std::uniform_real_distribution<double> ux(-1.0, 1.0);
  for (auto &v : x) v = ux(rng);
  double *xd = nullptr, *yd = nullptr;
  b2p::palace::Check(b2p_malloc(ctx, sizeof(double) * N, (void **)&xd), ctx);
  b2p::palace::Check(b2p_malloc(ctx, sizeof(double) * N, (void **)&yd), ctx);
  b2p::palace::Check(b2p_memcpy_h2d(ctx, xd, x.data(), sizeof(double) * N, nullptr), ctx);
  // (the mock Vector is a host array; the adapters' Read(true) pointers are passed through, so apply on device buffers directly)
  b2p_op *sub = op->SubOperators()[0];
  b2p::palace::Check(b2p_op_apply(sub, xd, yd, nullptr), ctx);
  b2p::palace::Check(b2p_memcpy_d2h(ctx, y.data(), yd, sizeof(double) * N, nullptr), ctx);
  b2p::palace::Check(b2p_memcpy_h2d(ctx, yd, std::vector<double>(N, 0.0).data(), sizeof(double) * N, nullptr), ctx);
  b2p::palace::Check(b2p_op_diag_add(sub, yd, nullptr), ctx);
  b2p::palace::Check(b2p_memcpy_d2h(ctx, d.data(), yd, sizeof(double) * N, nullptr), ctx);
  b2p::palace::Check(b2p_ctx_sync(ctx, nullptr), ctx);

  // ---- the reference-style path on the same data: dense tables in native order, native restriction with signs, q-data ----
  std::vector<double> qdata((size_t)ne * 11 * Q), xe((size_t)ne * 3 * 8);
  const int lex2nat[8] = {0, 1, 3, 2, 4, 5, 7, 6};
  for (int e = 0; e < ne; e++)
    for (int c = 0; c < 3; c++)
      for (int l = 0; l < 8; l++) xe[((size_t)e * 3 + c) * 8 + l] = nodes[c * nverts + node_fes.elem_dofs[e][lex2nat[l]]];
  orc_geom_hex_qdata(ne, 1, q1d, xe.data(), mesh.attributes.data(), qdata.data());
  std::vector<int> idx((size_t)ne * P);
  std::vector<signed char> ori((size_t)ne * P);
  for (int e = 0; e < ne; e++)
    for (int i = 0; i < P; i++)
    {
      const int dd = nd_fes.elem_dofs[e][i];
      idx[(size_t)e * P + i] = dd >= 0 ? dd : -1 - dd;
      ori[(size_t)e * P + i] = dd >= 0 ? 1 : -1;
    }
  std::vector<double> y_ref(N, 0.0), d_ref(N, 0.0);
  if (mode == "dense_co")
    orc_apply_add_co(kind, ne, P, Q, interp.data(), curl.data(), idx.data(), co_ref.data(), qdata.data(), blob.data(), x.data(), y_ref.data());
  else
    orc_apply_add(kind, ne, P, Q, interp.data(), curl.data(), idx.data(), ori.data(), qdata.data(), blob.data(), x.data(), y_ref.data());
  if (mode == "dense_co")
    d_ref = d;  // (the oracle has no diagonal through the tridiagonal restriction: tests/test_tet_gpu.py covers it)
  else
    orc_diag_add(kind, ne, P, Q, interp.data(), curl.data(), idx.data(), qdata.data(), blob.data(), d_ref.data());
  double en = 0, rn = 0, ed = 0, rd = 0;
  for (int i = 0; i < N; i++)
  {
    en += (y[i] - y_ref[i]) * (y[i] - y_ref[i]);
    rn += y_ref[i] * y_ref[i];
    ed += (d[i] - d_ref[i]) * (d[i] - d_ref[i]);
    rd += d_ref[i] * d_ref[i];
  }
  const double ea = std::sqrt(en / rn), edg = std::sqrt(ed / rd);
  op.reset();
  b2p_geom_destroy(geom);
  b2p_free(ctx, xd);
  b2p_free(ctx, yd);
  b2p_ctx_destroy(ctx);
  std::printf("GLUE_EXEC %s apply_err=%.3e diag_err=%.3e dofs=%d\n", (ea < 1e-12 && edg < 1e-12) ? "OK" : "FAIL", ea, edg, N);
  return (ea < 1e-12 && edg < 1e-12) ? 0 : 1;
}
