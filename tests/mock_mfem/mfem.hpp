// TEST INFRASTRUCTURE ONLY: the handful of MFEM declarations include/b2p_palace.hpp touches, so that the adapter header can
// at least be compiled (syntax + types against the C ABI) in a container without MFEM. Not a functional MFEM.
#pragma once
#include <cstdlib>
#include <iostream>
#include <vector>
#define MFEM_VERSION 40700
#define MFEM_ABORT(msg)                  \
  do                                     \
  {                                      \
    std::cerr << msg << std::endl;       \
    std::abort();                        \
  } while (0)
#define MFEM_VERIFY(c, msg) \
  do                        \
  {                         \
    if (!(c)) MFEM_ABORT(msg); \
  } while (0)
namespace mfem
{
class Vector
{
  double *d = nullptr;
  int n = 0;

public:
  Vector() = default;
  explicit Vector(int n_) : d(new double[n_]()), n(n_) {}
  int Size() const { return n; }
  const double *Read(bool = true) const { return d; }
  double *Write(bool = true) { return d; }
  double *ReadWrite(bool = true) { return d; }
  Vector &operator=(double v)
  {
    for (int i = 0; i < n; i++) d[i] = v;
    return *this;
  }
  void SetSize(int n_)
  {
    delete[] d;
    d = new double[n_]();
    n = n_;
  }
  double *HostWrite() { return d; }
  const double *HostRead() const { return d; }
  double &operator[](int i) { return d[i]; }
  const double &operator[](int i) const { return d[i]; }
};
template <typename T>
class Array
{
  T *d = nullptr;
  int n = 0;

public:
  Array() = default;
  Array(const Array &o) : d(o.n ? new T[o.n] : nullptr), n(o.n)
  {
    for (int i = 0; i < n; i++) d[i] = o.d[i];
  }
  int Size() const { return n; }
  const T *HostRead() const { return d; }
  const T &operator[](int i) const { return d[i]; }
  void Assign(const std::vector<T> &v)
  {
    delete[] d;
    n = (int)v.size();
    d = n ? new T[n] : nullptr;
    for (int i = 0; i < n; i++) d[i] = v[i];
  }
};
// ---- the MFEM objects the integrator glue of b2p_palace.hpp reads (fem/libceed/basis.cpp:15-85, restriction.cpp:113-297,
// fem/mesh.cpp:146-209). The mock holds plain arrays that the test driver fills; the member functions have MFEM's
// signatures and conventions (column-major DofToQuad tables, -1-d encoding of flipped dofs, byNODES vdofs). ----
struct IntegrationPoint
{
  double x = 0.0, weight = 0.0, y = 0.0, z = 0.0;
};
class IntegrationRule
{
public:
  std::vector<IntegrationPoint> pts;
  int GetNPoints() const { return (int)pts.size(); }
  const IntegrationPoint &IntPoint(int i) const { return pts[i]; }
};
struct DofToQuad
{
  enum Mode
  {
    FULL,
    TENSOR
  };
  int ndof = 0, nqpt = 0;
  std::vector<double> B, G;  // B[q + nqpt * d]
  // FULL mode of a vector element: column-major (ndof, nqpt, dim) = row-major [dim][nqpt][ndof], values and curls / gradients
  std::vector<double> Bt, Gt;
};
class FiniteElement
{
protected:
  int order;

public:
  explicit FiniteElement(int p) : order(p) {}
  virtual ~FiniteElement() = default;
  int GetOrder() const { return order; }
};
// any vector element seen through its full tables (what InitNonTensorBasis reads, fem/libceed/basis.cpp:40-85)
class VectorFiniteElement : public FiniteElement
{
public:
  DofToQuad full_maps;
  int dim = 3;
  explicit VectorFiniteElement(int p) : FiniteElement(p) {}
  int GetDim() const { return dim; }
  const DofToQuad &GetDofToQuad(const IntegrationRule &, DofToQuad::Mode) const { return full_maps; }
};
// element dof transformation of ND tets / prisms with p >= 2 (fem/doftrans.hpp): the mock holds the matrix the primal inverse applies
class DofTransformation
{
public:
  int n = 0;
  std::vector<double> Minv;  // row-major [n][n]
  void InvTransformPrimal(Vector &v) const
  {
    std::vector<double> t(n, 0.0);
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) t[i] += Minv[(size_t)i * n + j] * v[j];
    for (int i = 0; i < n; i++) v[i] = t[i];
  }
};
class DenseMatrix
{
public:
  double a[9] = {0};
  double &operator()(int i, int j) { return a[i + 3 * j]; }
  const double &operator()(int i, int j) const { return a[i + 3 * j]; }
  double Det() const
  {
    return a[0] * (a[4] * a[8] - a[7] * a[5]) - a[3] * (a[1] * a[8] - a[7] * a[2]) + a[6] * (a[1] * a[5] - a[4] * a[2]);
  }
};
class TensorBasisElement
{
public:
  Array<int> dof_map;  // lexicographic -> native (-1 - native for a sign flip)
  virtual ~TensorBasisElement() = default;
  const Array<int> &GetDofMap() const { return dof_map; }
};
// H1 hexahedron (nodal tensor element): one 1-D basis
class NodalTensorFiniteElement : public FiniteElement, public TensorBasisElement
{
public:
  DofToQuad maps;
  explicit NodalTensorFiniteElement(int p) : FiniteElement(p) {}
  const DofToQuad &GetDofToQuad(const IntegrationRule &, DofToQuad::Mode) const { return maps; }
};
// ND hexahedron: closed and open 1-D bases
class VectorTensorFiniteElement : public FiniteElement, public TensorBasisElement
{
public:
  DofToQuad closed_maps, open_maps;
  explicit VectorTensorFiniteElement(int p) : FiniteElement(p) {}
  const DofToQuad &GetDofToQuad(const IntegrationRule &, DofToQuad::Mode) const { return closed_maps; }
  const DofToQuad &GetDofToQuadOpen(const IntegrationRule &, DofToQuad::Mode) const { return open_maps; }
};
class FiniteElementSpace
{
public:
  const FiniteElement *fe = nullptr;
  int ne = 0, vsize = 0, vdim = 1;
  std::vector<std::vector<int>> elem_dofs;  // scalar dofs, -1 - d for a flipped one
  int GetNE() const { return ne; }
  int GetVSize() const { return vsize; }
  int GetVDim() const { return vdim; }
  const FiniteElement *GetFE(int) const { return fe; }
  void GetElementDofs(int e, Array<int> &dofs) const { dofs.Assign(elem_dofs[e]); }
  std::vector<const DofTransformation *> elem_trans;  // per element or empty
  const DofTransformation *GetElementDofs(int e, Array<int> &dofs, int) const
  {
    dofs.Assign(elem_dofs[e]);
    return elem_trans.empty() ? nullptr : elem_trans[e];
  }
  void GetElementVDofs(int e, Array<int> &vdofs) const  // Ordering::byNODES
  {
    std::vector<int> v;
    const int nd = vsize / vdim;
    for (int c = 0; c < vdim; c++)
      for (int d : elem_dofs[e]) v.push_back(d + c * nd);
    vdofs.Assign(v);
  }
};
class GridFunction : public Vector
{
  const FiniteElementSpace *fes;

public:
  GridFunction(const FiniteElementSpace *f, int n) : Vector(n), fes(f) {}
  const FiniteElementSpace *FESpace() const { return fes; }
  void GetSubVector(const Array<int> &dofs, Vector &out) const
  {
    out.SetSize(dofs.Size());
    for (int i = 0; i < dofs.Size(); i++) out.HostWrite()[i] = Read()[dofs[i]];
  }
};
// trilinear hexahedron through its 8 nodes (MFEM vertex order), enough for Jacobians at integration points
class ElementTransformation
{
public:
  double X[3][8];
  DenseMatrix J;
  void SetIntPoint(const IntegrationPoint *ip)
  {
    static const int cx[8] = {0, 1, 1, 0, 0, 1, 1, 0}, cy[8] = {0, 0, 1, 1, 0, 0, 1, 1}, cz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
    const double xi[3] = {ip->x, ip->y, ip->z};
    for (int c = 0; c < 3; c++)
      for (int d = 0; d < 3; d++) J(c, d) = 0.0;
    for (int v = 0; v < 8; v++)
    {
      const int cc[3] = {cx[v], cy[v], cz[v]};
      for (int d = 0; d < 3; d++)
      {
        double g = cc[d] ? 1.0 : -1.0;
        for (int o = 0; o < 3; o++)
          if (o != d) g *= cc[o] ? xi[o] : 1.0 - xi[o];
        for (int c = 0; c < 3; c++) J(c, d) += X[c][v] * g;
      }
    }
  }
  const DenseMatrix &Jacobian() const { return J; }
  double Weight() const { return J.Det(); }
};
class Mesh
{
public:
  int ne = 0;
  std::vector<int> attributes;
  const GridFunction *nodes = nullptr;
  mutable ElementTransformation trans;
  ElementTransformation *GetElementTransformation(int e) const
  {
    const FiniteElementSpace *nf = nodes->FESpace();
    const int nd = nf->vsize / nf->vdim;
    for (int c = 0; c < 3; c++)
      for (int v = 0; v < 8; v++) trans.X[c][v] = (*nodes)[nf->elem_dofs[e][v] + c * nd];
    return &trans;
  }
  int GetNE() const { return ne; }
  int GetAttribute(int e) const { return attributes[e]; }
  const GridFunction *GetNodes() const { return nodes; }
};

class Operator
{
protected:
  int height, width;

public:
  explicit Operator(int s = 0) : height(s), width(s) {}
  Operator(int h, int w) : height(h), width(w) {}
  virtual ~Operator() = default;
  int Height() const { return height; }
  int Width() const { return width; }
  virtual void Mult(const Vector &x, Vector &y) const = 0;
  virtual void MultTranspose(const Vector &, Vector &) const {}
  virtual void AddMult(const Vector &, Vector &, const double = 1.0) const {}
  virtual void AddMultTranspose(const Vector &, Vector &, const double = 1.0) const {}
  virtual void AssembleDiagonal(Vector &) const {}
};
class Solver : public Operator
{
public:
  explicit Solver(int s = 0) : Operator(s) {}
  virtual void SetOperator(const Operator &op) = 0;
};
}  // namespace mfem
