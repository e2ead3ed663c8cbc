// Drop-in layer, part 7b: EoS_Interpolator (reference include/eos_interpolator.h:61-163, src/eos_interpolator.cpp).
// The reference reads its pre-inverted tables from HDF5; the tables of this library are plain arrays. The constructor reads the
// flat binary file ccake_b200/eos_preinverted.py:save_binary writes ("CCBPRE01", format below) and hands the arrays to the library;
// fill_thermodynamics is phase 04 of the device run (ccb_update_thermodynamics).
//   file: char magic[8] = "CCBPRE01"; int32 flat; int32 ntiles;
//         flat = 0: per tile { int32 slot[4], N[4], log_spacing[4]; double min[4], step[4]; double vars[11 * N0*N1*N2*N3]; }
//         flat = 1: int32 Ns, NB; 12 arrays of Ns*NB doubles: s nB T muB muS muQ e p cs2 chiTT chiTB chiBB (MeV units)
#ifndef EOS_INTERPOLATOR_H
#define EOS_INTERPOLATOR_H

#include <cstring>
#include <filesystem>
#include <fstream>
#include <memory>
#include <stdexcept>
#include <vector>

#include "device_context.h"

namespace fs = std::filesystem;

namespace ccake {

class EoS_Interpolator {
 public:
  EoS_Interpolator() = delete;
  EoS_Interpolator(fs::path path_to_eos_table) {
    std::ifstream in(path_to_eos_table, std::ios::binary);
    if (!in) throw std::runtime_error("EoS table not found at " + path_to_eos_table.string());
    char magic[8];
    in.read(magic, 8);
    if (std::memcmp(magic, "CCBPRE01", 8) != 0) throw std::runtime_error("EoS_Interpolator: not a CCBPRE01 pre-inverted table: " + path_to_eos_table.string());
    int32_t ntiles = 0;
    in.read(reinterpret_cast<char*>(&flat_), 4);
    in.read(reinterpret_cast<char*>(&ntiles), 4);
    if (flat_) {
      in.read(reinterpret_cast<char*>(&Ns_), 4);
      in.read(reinterpret_cast<char*>(&NB_), 4);
      flat_arrays_.assign(12, std::vector<double>((size_t)Ns_ * NB_));
      for (auto& a : flat_arrays_) in.read(reinterpret_cast<char*>(a.data()), sizeof(double) * a.size());
    } else {
      tiles_.resize(ntiles);
      tile_vars_.resize(ntiles);
      for (int k = 0; k < ntiles; ++k) {
        ccb_preinverted_tile& t = tiles_[k];
        in.read(reinterpret_cast<char*>(t.slot), 16); in.read(reinterpret_cast<char*>(t.N), 16); in.read(reinterpret_cast<char*>(t.log_spacing), 16);
        in.read(reinterpret_cast<char*>(t.min), 32); in.read(reinterpret_cast<char*>(t.step), 32);
        tile_vars_[k].resize((size_t)11 * t.N[0] * t.N[1] * t.N[2] * t.N[3]);
        in.read(reinterpret_cast<char*>(tile_vars_[k].data()), sizeof(double) * tile_vars_[k].size());
        t.vars = tile_vars_[k].data();
      }
    }
    if (!in) throw std::runtime_error("EoS_Interpolator: truncated table file " + path_to_eos_table.string());
  }
  bool is_flat() const { return flat_ != 0; }
  // hand the table to the library (SPHWorkstation::initialize does this right after constructing the interpolator)
  void attach(std::shared_ptr<b200::Device> dev) {
    device = dev;
    if (flat_) {
      ccb_preinverted_flat f;
      f.Ns = Ns_; f.NB = NB_;
      const double** slot[12] = {&f.s, &f.nB, &f.T, &f.muB, &f.muS, &f.muQ, &f.e, &f.p, &f.cs2, &f.chiTT, &f.chiTB, &f.chiBB};
      for (int k = 0; k < 12; ++k) *slot[k] = flat_arrays_[k].data();
      device->check(ccb_set_preinverted_flat(device->ctx, &f));
    } else {
      device->check(ccb_set_preinverted_table(device->ctx, (int32_t)tiles_.size(), tiles_.data()));
    }
  }
  // phase 04 of the device run: EoS_Interpolator::fill_thermodynamics (src/eos_interpolator.cpp:531-627)
  template <class Particles> void fill_thermodynamics(Particles&, const double) { device->check(ccb_update_thermodynamics(device->ctx)); }

 private:
  std::shared_ptr<b200::Device> device;
  int32_t flat_ = 0, Ns_ = 0, NB_ = 0;
  std::vector<ccb_preinverted_tile> tiles_;
  std::vector<std::vector<double>> tile_vars_, flat_arrays_;
};

}  // namespace ccake
#endif
