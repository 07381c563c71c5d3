// Minimal stand-in for <Rcpp.h>: just enough surface to TYPE-CHECK r/src/gpu_glue.cpp against
// include/tpg_b200.h on a machine without R (tests/test_abi.py compiles the glue with
// `g++ -fsyntax-only -Itests/stubs`).  Nothing here runs; it only keeps the written-out R binding
//  from drifting away from the C ABI it calls.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace Rcpp {

template <typename T>
class Vector {
 public:
  Vector() {}
  explicit Vector(long n) : v_(size_t(n)) {}
  T* begin() { return v_.data(); }
  const T* begin() const { return v_.data(); }
  long size() const { return long(v_.size()); }
  Vector operator+(T x) const { Vector r(*this); for (auto& e : r.v_) e += x; return r; }
 private:
  std::vector<T> v_;
};
typedef Vector<int> IntegerVector;
typedef Vector<double> NumericVector;

class NumericMatrix {
 public:
  NumericMatrix(int nr, int nc) : nr_(nr), nc_(nc), v_(size_t(nr) * size_t(nc)) {}
  double* begin() { return v_.data(); }
  int nrow() const { return nr_; }
  int ncol() const { return nc_; }
 private:
  int nr_, nc_;
  std::vector<double> v_;
};

struct FieldProxy {  // what Environment::operator[] hands out
  operator NumericVector() const { return NumericVector(); }
};
template <typename T> T as(const FieldProxy&) { return T(); }

class Environment {
 public:
  FieldProxy operator[](const char*) const { return FieldProxy(); }
};

struct Named {
  template <typename T> Named& operator=(const T&) { return *this; }
};
struct NamedPlaceholder {
  Named operator[](const char*) const { return Named(); }
};
static const NamedPlaceholder _ = NamedPlaceholder();

class List {
 public:
  template <typename... A> static List create(const A&...) { return List(); }
  bool containsElementNamed(const char*) const { return false; }
  FieldProxy operator[](const char*) const { return FieldProxy(); }
};

[[noreturn]] inline void stop(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw std::runtime_error(buf);
}

}  // namespace Rcpp
