// Drop-in layer, part 7 of 8: EquationOfState and EoS_Interpolator facades (reference include/eos.h:62-210,
// include/eos_base.h:12-50, include/eos_interpolator.h:135-163) over the host-query entry points of the C ABI
// (ccb_eos_full_thermo, ccb_eos_locate, ccb_set_preinverted_*). The per-particle inversion of phase 04 does NOT go through
// these classes (it runs on the device inside ccb_step); they serve the host code that queries the EoS directly -- initial
// conditions, freeze-out energy, closure tests.
#ifndef EOS_H
#define EOS_H

#include <cmath>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

#include "device_context.h"
#include "constants.h"
#include "thermodynamic_info.h"

using namespace constants;   // as the reference's eos_table.h:14 / rootfinder.h:20 do: BSQHydro.cpp names hbarc_GeVfm unqualified

namespace ccake {

class EquationOfState {
 public:
  EquationOfState() {}
  ~EquationOfState() {}
  void set_SettingsPtr(std::shared_ptr<Settings> settingsPtr_in) { settingsPtr = settingsPtr_in; }
  // share the context of a run (SPHWorkstation::initialize does), or create one of its own in init()
  void attach(std::shared_ptr<b200::Device> dev) { device = dev; }
  // EquationOfState::init (src/eos_initialization.cpp:30-455): table load + fall-back chain, all inside the library
  void init() {
    if (!device) {
      device = std::make_shared<b200::Device>();
      ccb_config c = b200::to_config(*settingsPtr);
      c.dim = 2;
      device->create(c);
      if (settingsPtr->eos_type == "table")
        device->check(ccb_load_eos_table_dat(device->ctx, (settingsPtr->eos_path + "/thermo.dat").c_str(), settingsPtr->normalize_by_T ? 1 : 0));
    }
    default_eos_name = settingsPtr->eos_type;
    current_eos_name = default_eos_name;
  }
  // members of chosen_EOSs in the order of EquationOfState::set_up_chosen_EOSs (src/eos_initialization.cpp:249-455)
  int member_of(const std::string& name) const {
    if (settingsPtr->eos_type != "table") return 0;
    return name == "tanh_conformal" ? 1 : name == "conformal" ? 2 : name == "conformal_diagonal" ? 3 : 0;
  }
  std::string name_of(int member) const {
    static const char* chain[4] = {"table", "tanh_conformal", "conformal", "conformal_diagonal"};
    return settingsPtr->eos_type == "table" ? chain[member < 0 || member > 3 ? 0 : member] : settingsPtr->eos_type;
  }
  // set the phase-diagram point and evaluate the thermodynamics there (src/eos.cpp:128-221)
  void tbqs(double setT, double setmuB, double setmuQ, double setmuS, const std::string& eos_name) {
    current_eos_name = (eos_name == "default") ? default_eos_name : eos_name;
    const double pt[4] = {setT, setmuB, setmuQ, setmuS};
    device->check(ccb_eos_full_thermo(device->ctx, member_of(current_eos_name), 1, pt, full));
    point[0] = setT; point[1] = setmuB; point[2] = setmuS; point[3] = setmuQ;
    have16 = false;
  }
  void tbqs(std::vector<double>& tbqsIn, const std::string& eos_name) { tbqs(tbqsIn[0], tbqsIn[1], tbqsIn[2], tbqsIn[3], eos_name); }
  // (s, rhoB, rhoS, rhoQ) -> phase-diagram point, from the current point, through the fall-back chain (src/eos.cpp:574-804)
  bool update_s(double sin, double Bin, double Sin, double Qin, bool = false) { return locate(0, sin, Bin, Sin, Qin); }
  bool update_s(double sin) { return update_s(sin, 0.0, 0.0, 0.0); }
  // (e, rhoB, rhoS, rhoQ) -> s (src/eos.cpp:834-977)
  double s_out(double ein, double Bin, double Sin, double Qin, bool& solution_found, bool = false) {
    solution_found = locate(1, ein, Bin, Sin, Qin);
    return solution_found ? th16[5] : 0.0;
  }
  double s_out(double ein, bool& solution_found) { return s_out(ein, 0.0, 0.0, 0.0, solution_found); }
  // the same for n density vectors (e, rhoB, rhoS, rhoQ) at once: thermo[n][16] in CCB_F_THERMO order, eos[n] = answering member of
  // chosen_EOSs or -1 where s_out reports no solution. One device call (SPHWorkstation::freeze_in_match iterates over all cells).
  void s_out_batch(int64_t n, const double* dens, double* thermo, int32_t* eos) {
    if (n > 0) device->check(ccb_eos_locate(device->ctx, 1, n, dens, nullptr, nullptr, thermo, eos));
  }

  double T() const { return point[0]; }
  double muB() const { return point[1]; }
  double muS() const { return point[2]; }
  double muQ() const { return point[3]; }
  // order of the 17 fields: p s B S Q e cs2 chiBB ... (include/ccake_b200.h: ccb_eos_full_thermo)
  double p() const { return have16 ? th16[9] : full[0]; }
  double s() const { return have16 ? th16[5] : full[1]; }
  double B() const { return have16 ? th16[6] : full[2]; }
  double S() const { return have16 ? th16[7] : full[3]; }
  double Q() const { return have16 ? th16[8] : full[4]; }
  double e() const { return have16 ? th16[4] : full[5]; }
  double cs2() const { return have16 ? th16[10] : full[6]; }
  double w() const { return e() + p(); }
  std::string get_current_eos_name() const { return current_eos_name; }
  // thermodynamic_info of the located point (include/eos.h:121-148); the dalpha block is filled by the device run, not here
  void set_thermo(thermodynamic_info& thermo) {
    thermo.eos_name = current_eos_name;
    thermo.eos_type_is_table = (thermo.eos_name == "table") ? 1.0 : 0.0;
    thermo.T = T(); thermo.muB = muB(); thermo.muS = muS(); thermo.muQ = muQ();
    thermo.p = p(); thermo.s = s(); thermo.rhoB = B(); thermo.rhoS = S(); thermo.rhoQ = Q(); thermo.e = e(); thermo.w = w(); thermo.cs2 = cs2();
    if (have16) { thermo.dwds = th16[12]; thermo.dwdB = th16[13]; thermo.dwdS = th16[14]; thermo.dwdQ = th16[15]; }
  }
  // e(T_FO, mu = 0) of the default EoS (src/eos.cpp:553-557)
  double efreeze(double TFO) { tbqs(TFO, 0.0, 0.0, 0.0, default_eos_name); return e(); }
  double sfreeze(double TFO) { tbqs(TFO, 0.0, 0.0, 0.0, default_eos_name); return s(); }

 private:
  bool locate(int energy, double a, double Bin, double Sin, double Qin) {
    const double dens[4] = {a, Bin, Sin, Qin};
    const double seed[4] = {point[0], point[1], point[2], point[3]};
    const int32_t start = member_of(current_eos_name);
    int32_t eos = -1;
    device->check(ccb_eos_locate(device->ctx, energy, 1, dens, seed, &start, th16, &eos));
    if (eos < 0) return false;
    for (int k = 0; k < 4; ++k) point[k] = th16[k];
    current_eos_name = name_of(eos);
    have16 = true;
    return true;
  }
  std::shared_ptr<Settings> settingsPtr;
  std::shared_ptr<b200::Device> device;
  std::string default_eos_name, current_eos_name;
  double point[4] = {0, 0, 0, 0};  // T, muB, muS, muQ
  double full[17] = {0}, th16[16] = {0};
  bool have16 = false;
};

}  // namespace ccake
#endif
