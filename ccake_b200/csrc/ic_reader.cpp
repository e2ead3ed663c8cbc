// Initial-condition text readers at scale (host code of the library; no CUDA here).
// Reference: BSQHydro::read_ccake (src/BSQHydro.cpp:207-425) and BSQHydro<2,EoM_default>::read_ICCING (:84-145). Those parse with
// getline + istringstream into one ~5 KB Particle<D> per row, serially: minutes and tens of GB at 8-64 M rows. Here the file is
// read once, cut at line boundaries into one chunk per hardware thread, parsed with std::from_chars straight into the
// structure-of-arrays block ccb_upload_ic takes (ccb_ic), rows below the energy cut-off dropped as the reference drops them.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ccake_b200.h"

namespace {
constexpr double kHbarcGeVfm = 0.1973269804;  // reference include/constants.h:15

struct Rows {
  std::vector<double> r, e, rhoB, rhoS, rhoQ, u, bulk, shv6, q;
};
inline const char* skip_ws(const char* p, const char* end) { while (p < end && (*p == ' ' || *p == '\t' || *p == '\r')) ++p; return p; }
// one number; a leading '+' (which from_chars rejects) is accepted
inline bool number(const char*& p, const char* end, double& v) {
  p = skip_ws(p, end);
  if (p < end && *p == '+') ++p;
  auto res = std::from_chars(p, end, v);
  if (res.ec != std::errc()) return false;
  p = res.ptr;
  return true;
}
}  // namespace

struct ccb_ic_file {
  Rows rows;
  ccb_ic view;
  double header[6];   // stepx, stepy, stepEta, xmin, ymin, etamin
  std::string err;
};

extern "C" {

// format: 0 = ccake (17 columns, 29 with the initial diffusion currents), 1 = ICCING (x y e rhoB rhoS rhoQ; 2-D, u = 0)
int ccb_ic_read(const char* path, int format, int input_as_entropy, double e_cutoff_GeV, int baryon, int strange, int electric, int with_diffusion,
                ccb_ic_file** out) {
  if (!path || !out) return CCB_ERR_INVALID;
  *out = nullptr;
  FILE* f = std::fopen(path, "rb");
  if (!f) return CCB_ERR_INVALID;
  std::fseek(f, 0, SEEK_END);
  const long size = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  std::vector<char> buf((size_t)size + 1);
  const size_t got = std::fread(buf.data(), 1, (size_t)size, f);
  std::fclose(f);
  if ((long)got != size) return CCB_ERR_INVALID;
  buf[size] = '\n';
  const char *begin = buf.data(), *end = buf.data() + size;
  auto* icf = new ccb_ic_file();
  // header line (src/BSQHydro.cpp:240-252, :103-112): "# ignore stepx stepy stepEta ignore xmin ymin etamin" | "ignore stepx stepy ignore xmin ymin"
  const char* nl = static_cast<const char*>(std::memchr(begin, '\n', (size_t)(end - begin)));
  if (!nl) nl = end;
  {
    const char* p = begin;
    if (format == 0) { const char* hash = static_cast<const char*>(std::memchr(p, '#', (size_t)(nl - p))); if (hash) p = hash + 1; }
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int want = format == 0 ? 8 : 6;
    for (int k = 0; k < want; ++k) if (!number(p, nl, v[k])) { delete icf; return CCB_ERR_INVALID; }
    if (format == 0) { icf->header[0] = v[1]; icf->header[1] = v[2]; icf->header[2] = v[3]; icf->header[3] = v[5]; icf->header[4] = v[6]; icf->header[5] = v[7]; }
    else { icf->header[0] = v[1]; icf->header[1] = v[2]; icf->header[2] = 0.0; icf->header[3] = v[4]; icf->header[4] = v[5]; icf->header[5] = 0.0; }
  }
  const char* body = nl < end ? nl + 1 : end;
  const int nthreads = (int)std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
  std::vector<const char*> cut(nthreads + 1);
  cut[0] = body; cut[nthreads] = end;
  for (int k = 1; k < nthreads; ++k) {
    const char* p = body + (size_t)((end - body) * (double)k / nthreads);
    const char* q = static_cast<const char*>(std::memchr(p, '\n', (size_t)(end - p)));
    cut[k] = q ? q + 1 : end;
  }
  std::vector<Rows> part(nthreads);
  std::vector<int> bad(nthreads, 0);
  const int ncol = format == 0 ? (with_diffusion ? 29 : 17) : 6;
  const double cut_fm = e_cutoff_GeV / kHbarcGeVfm;
  auto work = [&](int k) {
    Rows& R = part[k];
    const char* p = cut[k];
    while (p < cut[k + 1]) {
      const char* le = static_cast<const char*>(std::memchr(p, '\n', (size_t)(cut[k + 1] - p)));
      if (!le) le = cut[k + 1];
      const char* q = skip_ws(p, le);
      if (q < le && *q != '#') {
        double v[29];
        bool ok = true;
        for (int c = 0; c < ncol && ok; ++c) ok = number(q, le, v[c]);
        if (!ok) { bad[k] = 1; return; }
        bool keep;
        const double a = v[format == 0 ? 3 : 2];
        if (format == 0 && input_as_entropy) {   // ideal pion-gas estimate of e from s (src/BSQHydro.cpp:395-399)
          const double C = 3.0 * M_PI * M_PI / 90.0;
          keep = (3. * C) * std::pow(1. / (4. * C), 4. / 3.) * std::pow(a, 4. / 3.) > cut_fm;
        } else keep = a / kHbarcGeVfm > cut_fm;
        if (keep) {
          if (format == 0) {
            R.r.insert(R.r.end(), {v[0], v[1], v[2]}); R.e.push_back(a);
            R.rhoB.push_back(baryon ? v[4] : 0.0); R.rhoS.push_back(strange ? v[5] : 0.0); R.rhoQ.push_back(electric ? v[6] : 0.0);
            R.u.insert(R.u.end(), {v[7], v[8], v[9]}); R.bulk.push_back(v[10]);
            R.shv6.insert(R.shv6.end(), v + 11, v + 17);
            if (with_diffusion) {
              const int on[3] = {baryon, strange, electric};
              for (int c = 0; c < 3; ++c) for (int m = 0; m < 4; ++m) R.q.push_back(on[c] ? v[17 + 4 * c + m] : 0.0);
            }
          } else {
            R.r.insert(R.r.end(), {v[0], v[1], 0.0}); R.e.push_back(a);
            R.rhoB.push_back(baryon ? v[3] : 0.0); R.rhoS.push_back(strange ? v[4] : 0.0); R.rhoQ.push_back(electric ? v[5] : 0.0);
            R.u.insert(R.u.end(), {0.0, 0.0, 0.0}); R.bulk.push_back(0.0);
            R.shv6.insert(R.shv6.end(), 6, 0.0);
          }
        }
      }
      p = le < cut[k + 1] ? le + 1 : cut[k + 1];
    }
  };
  std::vector<std::thread> th;
  for (int k = 0; k < nthreads; ++k) th.emplace_back(work, k);
  for (auto& t : th) t.join();
  for (int k = 0; k < nthreads; ++k) if (bad[k]) { delete icf; return CCB_ERR_INVALID; }
  Rows& R = icf->rows;
  for (int k = 0; k < nthreads; ++k) {   // file order is kept: particle ids are row numbers, as in the reference
    Rows& P = part[k];
    R.r.insert(R.r.end(), P.r.begin(), P.r.end()); R.e.insert(R.e.end(), P.e.begin(), P.e.end());
    R.rhoB.insert(R.rhoB.end(), P.rhoB.begin(), P.rhoB.end()); R.rhoS.insert(R.rhoS.end(), P.rhoS.begin(), P.rhoS.end());
    R.rhoQ.insert(R.rhoQ.end(), P.rhoQ.begin(), P.rhoQ.end()); R.u.insert(R.u.end(), P.u.begin(), P.u.end());
    R.bulk.insert(R.bulk.end(), P.bulk.begin(), P.bulk.end()); R.shv6.insert(R.shv6.end(), P.shv6.begin(), P.shv6.end());
    R.q.insert(R.q.end(), P.q.begin(), P.q.end());
    P = Rows();
  }
  ccb_ic& v = icf->view;
  std::memset(&v, 0, sizeof(v));
  v.n = (int64_t)R.e.size();
  v.r = R.r.data(); v.e_or_s = R.e.data(); v.rhoB = R.rhoB.data(); v.rhoS = R.rhoS.data(); v.rhoQ = R.rhoQ.data(); v.u = R.u.data();
  v.bulk = R.bulk.data(); v.shv6 = R.shv6.data();
  v.diffusion = (format == 0 && with_diffusion) ? R.q.data() : nullptr;
  *out = icf;
  return CCB_OK;
}
const ccb_ic* ccb_ic_file_view(const ccb_ic_file* f) { return f ? &f->view : nullptr; }
int ccb_ic_file_header(const ccb_ic_file* f, double header[6]) {
  if (!f || !header) return CCB_ERR_INVALID;
  for (int k = 0; k < 6; ++k) header[k] = f->header[k];
  return CCB_OK;
}
void ccb_ic_file_free(ccb_ic_file* f) { delete f; }

}  // extern "C"
