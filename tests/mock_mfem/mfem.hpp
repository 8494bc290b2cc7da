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
  double x = 0.0, weight = 0.0;
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
class Mesh
{
public:
  int ne = 0;
  std::vector<int> attributes;
  const GridFunction *nodes = nullptr;
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
