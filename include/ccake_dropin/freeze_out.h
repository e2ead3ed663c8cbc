// Drop-in layer, part 6 of 8: FreezeOut<D> (reference include/freeze_out.h:165-933). The status machine, the snapshots and the
// hypersurface interpolation run on the device (ccb_freeze_out); `results` is the host view of the records Output::print_freeze_out
// writes (src/output.cpp:417-600), one Row per particle with the reference's member names, refreshed by fetch_results().
#ifndef FREEZEOUT_H
#define FREEZEOUT_H

#include <memory>
#include <vector>

#include "system_state.h"

namespace ccake {

template <unsigned int D>
class FreezeOut {
 public:
  // ResultsTypes (include/freeze_out.h:35-49)
  struct Row {
    double shearsub[4][4], diffout[3][4], divT[3], divE[3], divP[3], rsub[3], uout[3];
    double divTtemp, divPpress, divEener, gsub, bulksub, thetasub, swsub, shear33sub, tlist, sFO, Efluc, Tfluc, muBfluc, muSfluc, muQfluc,
        cs2fzfluc, wfzfluc, rhoBfluc, rhoSfluc, rhoQfluc, eos_type;
    bool print;
  };
  std::vector<Row> results;   // freeze_results / results_h of the reference

  FreezeOut() = delete;
  FreezeOut(std::shared_ptr<Settings> settingsPtr_in, std::shared_ptr<SystemState<D>> systemPtr_in) : settingsPtr(settingsPtr_in), systemPtr(systemPtr_in) {}

  // check_freeze_out_status + bsqsvfreezeout of one time step (SPHWorkstation::freeze_out_particles, src/sph_workstation.cpp:1475-1485)
  void freeze_out_particles(int& count) {
    int64_t k = 0;
    systemPtr->device->check(ccb_freeze_out(systemPtr->device->ctx, &k));
    count = (int)k;
    systemPtr->number_part_fo += count;
  }
  // device -> `results` (what Cabana::deep_copy(results_h, results) does in Output::print_freeze_out)
  void fetch_results() {
    const size_t n = (size_t)systemPtr->n_particles;
    std::vector<double> raw(n * CCB_FO_ROW);
    std::vector<int32_t> pr(n);
    systemPtr->device->check(ccb_freeze_out_results(systemPtr->device->ctx, raw.data(), pr.data()));
    results.resize(n);
    for (size_t i = 0; i < n; ++i) {
      const double* r = raw.data() + i * CCB_FO_ROW;
      Row& o = results[i];
      for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) o.shearsub[a][b] = r[CCB_FO_SHEARSUB + 4 * a + b];
      for (int c = 0; c < 3; ++c) for (int m = 0; m < 4; ++m) o.diffout[c][m] = r[CCB_FO_DIFFOUT + 4 * c + m];
      for (int d = 0; d < 3; ++d) { o.divT[d] = r[CCB_FO_DIVT + d]; o.divE[d] = r[CCB_FO_DIVE + d]; o.divP[d] = r[CCB_FO_DIVP + d]; o.rsub[d] = r[CCB_FO_RSUB + d]; o.uout[d] = r[CCB_FO_UOUT + d]; }
      o.divTtemp = r[CCB_FO_DIVTTEMP]; o.divPpress = r[CCB_FO_DIVPPRESS]; o.divEener = r[CCB_FO_DIVEENER]; o.gsub = r[CCB_FO_GSUB];
      o.bulksub = r[CCB_FO_BULKSUB]; o.thetasub = r[CCB_FO_THETASUB]; o.swsub = r[CCB_FO_SWSUB]; o.shear33sub = r[CCB_FO_SHEAR33SUB];
      o.tlist = r[CCB_FO_TLIST]; o.sFO = r[CCB_FO_SFO]; o.Efluc = r[CCB_FO_EFLUC]; o.Tfluc = r[CCB_FO_TFLUC]; o.muBfluc = r[CCB_FO_MUBFLUC];
      o.muSfluc = r[CCB_FO_MUSFLUC]; o.muQfluc = r[CCB_FO_MUQFLUC]; o.cs2fzfluc = r[CCB_FO_CS2FZFLUC]; o.wfzfluc = r[CCB_FO_WFZFLUC];
      o.rhoBfluc = r[CCB_FO_RHOBFLUC]; o.rhoSfluc = r[CCB_FO_RHOSFLUC]; o.rhoQfluc = r[CCB_FO_RHOQFLUC]; o.eos_type = r[CCB_FO_EOS_TYPE];
      o.print = pr[i] != 0;
    }
  }

 private:
  std::shared_ptr<Settings> settingsPtr;
  std::shared_ptr<SystemState<D>> systemPtr;
};

}  // namespace ccake
#endif
