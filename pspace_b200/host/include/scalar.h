// scalar.h — the one compile-time switch of the library: `scalar` is double, or std::complex<double>
// when USE_COMPLEX is defined (same switch and helper names as the reference's src/include/scalar.h:12-55,
// so callers written against it compile unchanged).  The switch must be identical for libpspace.so and
// everything built on it.
#ifndef PSPACE_B200_SCALAR_H
#define PSPACE_B200_SCALAR_H

#include <complex>
#include <cstdlib>

using namespace std;   // the reference header exports namespace std to its includers; callers rely on it

typedef std::complex<double> Complex;
typedef double Real;

#ifdef USE_COMPLEX
typedef Complex scalar;
#define PSPACE_SCALAR_WIDTH 2
#define MPI_TYPE MPI_DOUBLE_COMPLEX
#else
typedef double scalar;
#define PSPACE_SCALAR_WIDTH 1
#define MPI_TYPE MPI_DOUBLE
#endif

inline double RealPart(const double &r) { return r; }
inline double RealPart(const std::complex<double> &c) { return c.real(); }
inline double ImagPart(const std::complex<double> &c) { return c.imag(); }
inline std::complex<double> dabs(const std::complex<double> &c) { return c.real() < 0.0 ? -c : c; }

#endif
