// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
// In-memory stand-in for the subset of HDF5 C++ the reference's EoS_Interpolator reads with (src/eos_interpolator.cpp:16-340):
// "files" are registered by name through h5shim::registry() (oracle/ref_driver.cpp: ccake_ref_h5_put_*); a file holds
// double / int datasets addressed by "group/name" paths and scalar attributes addressed by "group@name". Nothing touches a
// disk: the reference only checks std::filesystem::exists(path) itself, so the driver drops an empty placeholder file there.
// The hot-path EoS tables of the other oracle builds are read through the reference's TEXT loader (interpolatorND.cpp:125-190).
#ifndef CCAKE_ORACLE_H5CPP_SHIM_H
#define CCAKE_ORACLE_H5CPP_SHIM_H
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>
typedef unsigned long long hsize_t;
typedef long hid_t;
#define H5F_ACC_RDONLY 0u
#define H5P_DEFAULT 0
namespace h5shim {
struct File {
  std::map<std::string, std::vector<double>> dsets;      // "group/name" or "name"
  std::map<std::string, std::vector<hsize_t>> dims;
  std::map<std::string, double> attrs;                   // "group@name"
};
inline std::map<std::string, std::shared_ptr<File>>& registry() { static std::map<std::string, std::shared_ptr<File>> r; return r; }
inline std::map<hid_t, std::shared_ptr<File>>& open_files() { static std::map<hid_t, std::shared_ptr<File>> r; return r; }
[[noreturn]] inline void fail(const std::string& what) { std::fprintf(stderr, "oracle H5 shim: %s\n", what.c_str()); std::abort(); }
}  // namespace h5shim
namespace H5 {
struct Exception { const char* getCDetailMsg() const { return "oracle H5 shim"; } };
struct FileIException : Exception {};
struct PredType { static const int NATIVE_INT = 0; static const int NATIVE_DOUBLE = 1; };
struct DataSpace {
  std::vector<hsize_t> d;
  DataSpace() {}
  DataSpace(int) {}
  int getSimpleExtentDims(hsize_t* out, hsize_t*) const { for (size_t i = 0; i < d.size(); ++i) out[i] = d[i]; return (int)d.size(); }
};
struct Attribute {
  double v = 0.0;
  void read(int type, void* out) const { if (type == PredType::NATIVE_INT) *static_cast<int*>(out) = (int)v; else *static_cast<double*>(out) = v; }
};
struct DataSet {
  const std::vector<double>* data = nullptr;
  std::vector<hsize_t> dims;
  DataSpace getSpace() const { DataSpace s; s.d = dims; return s; }
  void read(void* out, int type) const {
    if (type == PredType::NATIVE_INT) { int* o = static_cast<int*>(out); for (size_t i = 0; i < data->size(); ++i) o[i] = (int)(*data)[i]; }
    else std::memcpy(out, data->data(), sizeof(double) * data->size());
  }
};
struct Group {
  std::shared_ptr<h5shim::File> f;
  std::string name;
  Attribute openAttribute(const std::string& a) const {
    auto it = f->attrs.find(name + "@" + a);
    if (it == f->attrs.end()) h5shim::fail("no attribute " + name + "@" + a);
    Attribute r; r.v = it->second; return r;
  }
  void close() {}
};
struct H5File {
  std::shared_ptr<h5shim::File> f;
  hid_t id = 0;
  H5File() {}
  H5File(const std::string& path, unsigned) {
    auto it = h5shim::registry().find(path);
    if (it == h5shim::registry().end()) h5shim::fail("file not registered: " + path);
    f = it->second;
    id = (hid_t)(h5shim::open_files().size() + 1);
    h5shim::open_files()[id] = f;
  }
  DataSet openDataSet(const std::string& n) const {
    auto it = f->dsets.find(n);
    if (it == f->dsets.end()) h5shim::fail("no dataset " + n);
    DataSet d; d.data = &it->second; d.dims = f->dims[n]; return d;
  }
  Group openGroup(const std::string& n) const { Group g; g.f = f; g.name = n; return g; }
  hid_t getId() const { return id; }
  void close() {}
};
}  // namespace H5
inline int H5Lexists(hid_t id, const char* name, int) {
  auto it = h5shim::open_files().find(id);
  return (it != h5shim::open_files().end() && it->second->dsets.count(name)) ? 1 : 0;
}
#endif
