// TEST INFRASTRUCTURE ONLY: the handful of MFEM declarations include/b2p_palace.hpp touches, so that the adapter header can
// at least be compiled (syntax + types against the C ABI) in a container without MFEM. Not a functional MFEM.
#pragma once
#include <cstdlib>
#include <iostream>
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
