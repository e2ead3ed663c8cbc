// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
// Declaration-level stand-in for HDF5 C++: the oracle build reads EoS tables
// through the reference's text loader (interpolatorND.cpp:125-190), never HDF5.
#ifndef CCAKE_ORACLE_H5CPP_SHIM_H
#define CCAKE_ORACLE_H5CPP_SHIM_H
#include <cstdio>
#include <cstdlib>
#include <string>
typedef unsigned long long hsize_t;
typedef long hid_t;
#define H5F_ACC_RDONLY 0u
#define H5P_DEFAULT 0
namespace H5 {
[[noreturn]] inline void shim_abort() { std::fprintf(stderr, "HDF5 is not available in the oracle shim\n"); std::abort(); }
struct Exception {};
struct FileIException : Exception {};
struct PredType { static const int NATIVE_INT = 0; static const int NATIVE_DOUBLE = 1; };
struct DataSpace {
  DataSpace() {}
  DataSpace(int) {}
  int getSimpleExtentDims(hsize_t*, hsize_t*) const { shim_abort(); }
};
struct Attribute { template <class T> void read(int, T*) const { shim_abort(); } };
struct DataSet {
  DataSpace getSpace() const { shim_abort(); }
  template <class T> void read(T*, int) const { shim_abort(); }
};
struct Group { Attribute openAttribute(const std::string&) const { shim_abort(); } };
struct H5File {
  H5File() {}
  H5File(const std::string&, unsigned) { shim_abort(); }
  DataSet openDataSet(const std::string&) const { shim_abort(); }
  Group openGroup(const std::string&) const { shim_abort(); }
  hid_t getId() const { return 0; }
};
}
inline int H5Lexists(hid_t, const char*, int) { return 0; }
#endif
